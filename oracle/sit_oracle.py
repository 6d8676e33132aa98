"""TEST INFRASTRUCTURE (oracle) -- self-contained fp32 restatement of the reference SiT and MPP wrappers.

Follows /root/reference/models/sit.py:25-82 (SiT) and /root/reference/models/mpp.py:25-134
(masked_patch_pretraining).  It does not import /root/reference, so it also runs on the GPU box.
``tests/test_oracle.py`` pins it against the reference modules themselves (when /root/reference is present)
and against the golden vectors in tests/golden generated from them.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .vit_shim import Transformer


class _Rearrange(nn.Module):
    """'b c n v -> b n (v c)'   (models/sit.py:49) -- parameter-free, keeps index 0 of the Sequential."""

    def forward(self, x):
        b, c, n, v = x.shape
        return x.permute(0, 2, 3, 1).reshape(b, n, v * c)


class OracleSiT(nn.Module):
    # models/sit.py:26-64
    def __init__(self, *, dim, depth, heads, mlp_dim, pool="cls", num_patches=20, num_classes=1, num_channels=4,
                 num_vertices=2145, dim_head=64, dropout=0.0, emb_dropout=0.0):
        super().__init__()
        assert pool in {"cls", "mean"}, "pool type must be either cls (cls token) or mean (mean pooling)"
        patch_dim = num_channels * num_vertices
        self.to_patch_embedding = nn.Sequential(_Rearrange(), nn.Linear(patch_dim, dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))

    # models/sit.py:66-82
    def forward(self, img):
        x = self.to_patch_embedding(img)
        b, n, _ = x.shape
        cls_tokens = self.cls_token.expand(b, -1, -1)
        x = torch.cat((cls_tokens, x), dim=1)
        x = x + self.pos_embedding[:, : (n + 1)]
        x = self.dropout(x)
        x = self.transformer(x)
        x = x.mean(dim=1) if self.pool == "mean" else x[:, 0]
        x = self.to_latent(x)
        return self.mlp_head(x)

    def encode(self, img):
        """Encoder output (B, T, D) -- not in the reference API; used by parity tests."""
        x = self.to_patch_embedding(img)
        b, n, _ = x.shape
        x = torch.cat((self.cls_token.expand(b, -1, -1), x), dim=1) + self.pos_embedding[:, : (n + 1)]
        return self.transformer(self.dropout(x))


def get_mask_from_prob(inputs, prob):
    # models/mpp.py:25-39
    batch, seq_len, _ = inputs.shape
    device = inputs.device
    max_masked = math.ceil(prob * seq_len)
    rand = torch.rand((batch, seq_len), device=device)
    _, sampled_indices = rand.topk(max_masked, dim=-1)
    new_mask = torch.zeros((batch, seq_len), device=device)
    new_mask.scatter_(1, sampled_indices, 1)
    return new_mask.bool()


def prob_mask_like(inputs, prob):
    # models/mpp.py:41-43 -- NOTE: drawn on the CPU generator
    batch, seq_length, _ = inputs.shape
    return torch.zeros((batch, seq_length)).float().uniform_(0, 1) < prob


def draw_mpp_masks(batch_bnk, mask_prob, replace_prob, swap_prob):
    """RNG call order of models/mpp.py:85-109.  Returns (mask, swap_sel, swap_src, replace_sel)."""
    mask = get_mask_from_prob(batch_bnk, mask_prob)                                      # :85
    swap_sel = None
    swap_src = None
    if swap_prob > 0:
        p = swap_prob / (1 - replace_prob)                                                 # :91
        rp = prob_mask_like(batch_bnk, p).to(mask.device)                                # :94
        swap_sel = mask * (rp == True)                                                   # :97  # noqa: E712
        swap_src = torch.randint(0, batch_bnk.shape[1], (batch_bnk.shape[0], batch_bnk.shape[1]),
                                 device=batch_bnk.device)                               # :99
    tm = prob_mask_like(batch_bnk, replace_prob).to(mask.device)                         # :109
    replace_sel = (mask * tm) == True                                                    # :111  # noqa: E712
    return mask, swap_sel, swap_src, replace_sel


class OracleMPP(nn.Module):
    # models/mpp.py:46-74
    def __init__(self, transformer, dim_in, dim_out, device, mask_prob=0.15, replace_prob=0.5, swap_prob=0.3,
                 channels=4, num_vertices=561):
        super().__init__()
        self.transformer = transformer
        self.dim_out = dim_out
        self.dim_in = dim_in
        self.to_original = nn.Linear(dim_in, dim_out)
        self.to_original.to(device)
        self.mask_prob = mask_prob
        self.replace_prob = replace_prob
        self.swap_prob = swap_prob
        self.mask_token = nn.Parameter(torch.randn(1, 1, channels * num_vertices))

    # models/mpp.py:77-134
    def forward(self, batch, masks=None):
        t = self.transformer
        b, c, n, v = batch.shape
        batch = batch.permute(0, 2, 3, 1).reshape(b, n, v * c)                           # :82
        if masks is None:
            masks = draw_mpp_masks(batch, self.mask_prob, self.replace_prob, self.swap_prob)
        mask, swap_sel, swap_src, replace_sel = masks
        corrupted = batch.clone().detach()                                               # :87
        if self.swap_prob > 0:
            randomized = corrupted[torch.arange(b).unsqueeze(-1), swap_src]              # :104
            corrupted[swap_sel] = randomized[swap_sel]                                   # :107
        corrupted[replace_sel] = self.mask_token.to(mask.device)                         # :112
        x = t.to_patch_embedding[-1](corrupted)                                          # :115
        x = torch.cat((t.cls_token.expand(b, -1, -1), x), dim=1)                         # :119-121
        x = x + t.pos_embedding[:, : (n + 1)]                                            # :124
        x = t.dropout(x)                                                                 # :125
        out = t.transformer(x)                                                           # :128
        out = self.to_original(out[:, 1:, :])                                            # :129
        loss = F.mse_loss(out[mask], batch[mask])                                        # :132
        return loss, out
