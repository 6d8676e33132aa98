"""TEST INFRASTRUCTURE (oracle) -- restatement of ``vit_pytorch.vit.Transformer`` (third-party, absent).

The reference imports it at /root/reference/models/sit.py:23, constructs it at sit.py:57 and calls it at
sit.py:76 and models/mpp.py:128.  ``vit-pytorch`` is un-pinned in requirements.txt:5 and not installable here
(no network), so this file restates the published algorithm of the 0.x series whose parameter layout is the
one pinned by /root/reference/utils/utils.py:18-33:

    transformer.layers.{i}.0.norm.{weight,bias}          PreNorm(LayerNorm) around Attention
    transformer.layers.{i}.0.fn.to_qkv.weight            Linear(dim, 3*heads*dim_head, bias=False)
    transformer.layers.{i}.0.fn.to_out.0.{weight,bias}   Linear(heads*dim_head, dim) (+ Dropout)
    transformer.layers.{i}.1.norm.{weight,bias}          PreNorm(LayerNorm) around FeedForward
    transformer.layers.{i}.1.fn.net.0.{weight,bias}      Linear(dim, mlp_dim); net.1 = GELU (erf); net.2 = Dropout
    transformer.layers.{i}.1.fn.net.3.{weight,bias}      Linear(mlp_dim, dim); net.4 = Dropout

No final LayerNorm (that key does not exist in utils.py:13-33; the head's LayerNorm is ``mlp_head.0``).
"""
import sys
import types

import torch
from torch import nn


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(self.norm(x), **kwargs)


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(
            nn.Linear(dim, hidden_dim),
            nn.GELU(),
            nn.Dropout(dropout),
            nn.Linear(hidden_dim, dim),
            nn.Dropout(dropout),
        )

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner_dim = dim_head * heads
        project_out = not (heads == 1 and dim_head == dim)
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = (
            nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout)) if project_out else nn.Identity()
        )

    def forward(self, x):
        b, n, _ = x.shape
        h = self.heads
        qkv = self.to_qkv(x).chunk(3, dim=-1)
        # 'b n (h d) -> b h n d'
        q, k, v = (t.reshape(b, n, h, -1).permute(0, 2, 1, 3) for t in qkv)
        dots = torch.matmul(q, k.transpose(-1, -2)) * self.scale
        attn = self.attend(dots)
        out = torch.matmul(attn, v)
        # 'b h n d -> b n (h d)'
        out = out.permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out(out)


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(
                nn.ModuleList(
                    [
                        PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                        PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout)),
                    ]
                )
            )

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x) + x
            x = ff(x) + x
        return x


def install_as_vit_pytorch():
    """Registers this module as ``vit_pytorch.vit`` so the reference's models/sit.py imports unmodified."""
    if "vit_pytorch.vit" in sys.modules:
        return
    pkg = types.ModuleType("vit_pytorch")
    pkg.__path__ = []
    sub = types.ModuleType("vit_pytorch.vit")
    sub.Transformer = Transformer
    sub.Attention = Attention
    sub.FeedForward = FeedForward
    sub.PreNorm = PreNorm
    pkg.vit = sub
    sys.modules["vit_pytorch"] = pkg
    sys.modules["vit_pytorch.vit"] = sub
