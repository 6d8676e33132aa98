"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the dropout keep-mask generator of the B200 path.

The reference drops activations with ``torch.nn.Dropout`` at four sites (models/sit.py:55,74 ``emb_dropout``; in every
encoder block the Dropout of ``Attention.to_out`` and the two Dropouts of ``FeedForward.net`` -- the vit-pytorch layout
pinned by utils/utils.py:18-33, restated in oracle/vit_shim.py).  Two dropout implementations never share a random
stream, so parity is anchored differently: the CUDA path defines its keep decisions as a pure function of
(seed, offset, site, element index) built on Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as
easy as 1, 2, 3", SC'11 -- a published algorithm, pinned here by the Random123 known-answer vectors), this file
restates that function in numpy, and ``install`` swaps the oracle's ``nn.Dropout`` modules for ones that apply exactly
those masks.  With that the oracle and the CUDA path compute the same function and are compared at the usual tolerances;
the mask generator itself is compared bit for bit (include/svit_b200.h: svit_dropout_mask).

Only tests/ may import this module.
"""
import numpy as np
import torch
import torch.nn as nn

SITE_TO_OUT, SITE_FF_ACT, SITE_FF_OUT = 0, 1, 2
SITE_EMB = 0xFFFF0000
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds; counters are uint32 arrays (or scalars), the key two python ints."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def keep_mask(n, p, seed, offset, site):
    """bool[n]: element i is kept iff word (i & 3) of Philox(counter = {i >> 2, site, offset_lo, offset_hi},
    key = {seed_lo, seed_hi}) >= floor(p * 2**32)."""
    groups = (n + 3) // 4
    q = np.arange(groups, dtype=np.uint64)
    words = philox4x32_10(q, np.uint64(site), np.uint64(offset & 0xFFFFFFFF), np.uint64(offset >> 32),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    r = np.stack(words, axis=1).reshape(-1)[:n]
    thresh = np.uint32(int(float(np.float32(p)) * 4294967296.0))
    return r >= thresh


class MaskedDropout(nn.Module):
    """nn.Dropout with the keep decisions of the B200 path (same ``training`` semantics, same 1/(1-p) scaling)."""

    def __init__(self, p, seed, offset, site):
        super().__init__()
        self.p, self.seed, self.offset, self.site = float(p), int(seed), int(offset), int(site)

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        keep = keep_mask(x.numel(), self.p, self.seed, self.offset, self.site)
        m = torch.from_numpy(keep).to(x.device).view(x.shape).to(x.dtype)
        scale = float(np.float32(1.0 / (1.0 - float(np.float32(self.p)))))
        return x * (m * scale)


def install(model, p, emb_p, seed, offset):
    """Replaces the four kinds of nn.Dropout of an oracle / reference SiT (attribute names of models/sit.py and the
    vit-pytorch layout) by MaskedDropout with the site numbering of include/svit_b200.h."""
    model.dropout = MaskedDropout(emb_p, seed, offset, SITE_EMB)
    for i, (attn, ff) in enumerate(model.transformer.layers):
        attn.fn.to_out[1] = MaskedDropout(p, seed, offset, 4 * i + SITE_TO_OUT)
        ff.fn.net[2] = MaskedDropout(p, seed, offset, 4 * i + SITE_FF_ACT)
        ff.fn.net[4] = MaskedDropout(p, seed, offset, 4 * i + SITE_FF_OUT)
    return model
