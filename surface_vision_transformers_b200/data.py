"""Data formats on either side of the hot path.

* ``preprocess_meshes`` -- the on-device equivalent of tools/preprocessing.py:62-84: per-channel z-score of the raw
  ico-6 hemisphere meshes, patch gather with the reference's index table, left hemispheres first then right.  The
  z-score runs in float64 like the reference's numpy code and is cast to float32 before the (bit-exact) gather, so
  the result equals ``np.load('{split}_data.npy').astype(np.float32)`` of the reference bit for bit.
* ``PatchedNpyDataset`` -- reader of the reference's ``{split}_data.npy`` (float64, (2S, C, N, V)) /
  ``{split}_labels.npy`` files (tools/train.py:97-113): casts to float32 once (``torch.from_numpy(..).float()``,
  train.py:107), keeps the arrays in PINNED host memory and hands out mini-batches as pinned tuples for
  ``DevicePrefetcher`` -- no DataLoader workers, no per-iteration pageable copies.
"""
import os

import numpy as np
import torch

from .gather import gather_patches, preprocessing_layout

__all__ = ["preprocess_meshes", "PatchedNpyDataset", "shard_order"]


def shard_order(n, shuffle, generator, rank, world, equal_shards):
    """Sample indices of rank ``rank`` for one pass over ``n`` samples (host logic, no CUDA)."""
    order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
    if equal_shards and world > 1 and n % world != 0 and n > 0:
        pad = world - n % world
        reps = (pad + n - 1) // n
        order = torch.cat([order] + [order] * reps)[: n + pad]       # wrap around, like DistributedSampler
    return order[rank::world]


def preprocess_meshes(hemis, means, stds, table):
    """hemis: (2S, C, 40962) CUDA tensor ordered L0, R0, L1, R1, ... (preprocessing.py:62-67), any float dtype;
    means / stds: (C,) ; table: (V, N) int32 -> (2S, C, N, V) float32, rows [0, S) left, [S, 2S) right."""
    if not hemis.is_cuda:
        raise RuntimeError("preprocess_meshes needs CUDA tensors (no CPU fallback)")
    C = hemis.shape[1]
    m = torch.as_tensor(means, dtype=torch.float64, device=hemis.device).reshape(1, C, 1)
    s = torch.as_tensor(stds, dtype=torch.float64, device=hemis.device).reshape(1, C, 1)
    normalised = ((hemis.double() - m) / s).float()          # preprocessing.py:72 in float64, train.py:107 cast
    return preprocessing_layout(gather_patches(normalised, table))


class PatchedNpyDataset:
    """``{split}_data.npy`` + ``{split}_labels.npy`` of the reference, pinned in host memory as float32."""

    def __init__(self, data_path, split, pin=True, stage_dtype=torch.float32):
        """``stage_dtype=torch.bfloat16`` keeps the (pinned) samples as bf16: the B200 ``SiT`` takes bf16 batches as they
        are and gives bit-identical results (its patch-embedding GEMM consumes bf16 either way), at half the
        host-to-device bytes per step.  Not for ``masked_patch_pretraining``, whose loss compares with the fp32 input."""
        if stage_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("stage_dtype must be torch.float32 or torch.bfloat16")
        data = np.load(os.path.join(data_path, f"{split}_data.npy"))
        labels = np.load(os.path.join(data_path, f"{split}_labels.npy"))
        if data.ndim != 4 or labels.shape[0] != data.shape[0]:
            raise ValueError(f"unexpected shapes {data.shape} / {labels.shape} (want (2S, C, N, V) and (2S,))")
        self.data = torch.from_numpy(data).float().to(stage_dtype)
        self.labels = torch.from_numpy(labels).float()
        if pin and torch.cuda.is_available():
            self.data = self.data.pin_memory()
            self.labels = self.labels.pin_memory()
        self._stage = None

    def __len__(self):
        return self.data.shape[0]

    @property
    def shape(self):
        return tuple(self.data.shape)

    def batches(self, batch_size, shuffle=False, generator=None, drop_last=False, rank=0, world=1, equal_shards=None):
        """Yields (x, y) host batches.  With shuffle, a permutation drawn from ``generator`` (same on every rank) is
        split into per-rank strided slices, so the ranks of a data-parallel job see disjoint samples.

        ``equal_shards`` (default: ``shuffle``, i.e. on for training passes): the permutation is padded by wrapping
        around to a multiple of ``world`` (as torch's DistributedSampler does), so that every rank yields the SAME
        number of batches with the SAME sizes.  A training pass needs that: every batch ends in gradient all-reduces,
        and a rank with one batch more (or a ragged last batch of another size) would hang or bias the average.
        Evaluation passes (no per-batch collective) keep the exact, un-padded split."""
        order = shard_order(len(self), shuffle, generator, rank, world, shuffle if equal_shards is None else equal_shards)
        pinned = self.data.is_pinned()
        for i in range(0, order.numel(), batch_size):
            idx = order[i:i + batch_size]
            if drop_last and idx.numel() < batch_size:
                return
            if not shuffle and world == 1:
                yield self.data[i:i + idx.numel()], self.labels[i:i + idx.numel()]   # contiguous views stay pinned
                continue
            x = torch.empty((idx.numel(),) + self.data.shape[1:], dtype=self.data.dtype, pin_memory=pinned)
            y = torch.empty((idx.numel(),), dtype=torch.float32, pin_memory=pinned)
            torch.index_select(self.data, 0, idx, out=x)
            torch.index_select(self.labels, 0, idx, out=y)
            yield x, y
