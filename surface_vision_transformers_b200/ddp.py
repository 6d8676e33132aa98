"""Batch-sharded data parallelism for the fused SiT path: one process per GPU, replicated weights,
ONE exchange step per iteration -- an averaging all-reduce of the flat gradient buffer (``ReduceOp.AVG`` inside NCCL, no
scaling launches), issued at the engine's communication windows (under the attention-backward kernels), in one piece after
backward, or range by range as the gradients become final (see ``DataParallel``).

The reference has no distributed code (single ``cuda:{gpu}``, tools/train.py:72); this is the one strategy the
north star adds.  Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests of the bucketing
logic, which use plain tensors instead of the engine).
"""
import torch
import torch.distributed as dist

__all__ = ["DataParallel", "FlatGradReducer"]

DEFAULT_OVERLAP = "window"


class FlatGradReducer:
    """All-reduces ranges of a flat gradient tensor asynchronously and finalises them (wait + average)."""

    def __init__(self, process_group=None, average=True):
        self.pg = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.pending = []
        # NCCL averages inside the collective (ReduceOp.AVG): no separate scaling launches after the all-reduce.
        # Other backends (gloo in the CPU tests) sum and scale.
        self.native_avg = average and dist.is_initialized() and dist.get_backend(process_group) == "nccl"

    def reduce_range(self, G, start, numel):
        if self.world == 1 or numel == 0:
            return
        seg = G[start:start + numel]
        op = dist.ReduceOp.AVG if self.native_avg else dist.ReduceOp.SUM
        work = dist.all_reduce(seg, op=op, group=self.pg, async_op=True)
        self.pending.append((work, seg))

    def finish(self):
        for work, seg in self.pending:
            work.wait()
            if self.average and not self.native_avg:
                seg.mul_(1.0 / self.world)
        self.pending = []


STAGE_WINDOW = 1000   # include/svit_b200.h: SVIT_STAGE_WINDOW


class DataParallel(torch.nn.Module):
    """Wraps a B200 ``SiT`` or ``masked_patch_pretraining``: forwards calls unchanged and installs the gradient
    hooks that overlap the flat-buffer all-reduce with backward.  ``broadcast_parameters`` syncs the replicas once."""

    def __init__(self, module, process_group=None, broadcast_parameters=True, overlap=None):
        """``overlap`` selects when the averaging all-reduce of the flat gradient buffer is issued:

        * ``"window"`` -- the ranges that are final are all-reduced at the engine's communication windows: right
          before the attention backward of a layer is enqueued (``SVIT_STAGE_WINDOW``), so the NCCL kernel starts together
          with that ~270 us one-CTA-per-(sample, head) kernel, which hands SMs out CTA by CTA and shares the chip gracefully;
          the last range (layer 0 + patch embedding) follows backward.
        * ``False`` / ``"none"`` -- ONE all-reduce of the whole flat buffer after backward.
        * ``True`` / ``"range"`` -- every stage's range as soon as its gradients are final, under whatever backward
          kernel runs next.
        The SVIT_DDP_OVERLAP environment variable (0 / 1 / window) overrides the default.

        Why the plain range-wise overlap loses: the GEMMs of the backward pass are persistent kernels with a static tile
        schedule over all 148 SMs.  An NCCL kernel that occupies a few SMs for the duration of a range's all-reduce keeps
        the GEMM CTAs assigned to those SMs from starting, and a kernel whose tiles were dealt out round-robin then waits
        for them -- measured on 8 B200s (SiT-small, batch 256 per GPU, same box): 20.77 ms single GPU, 21.77 ms with the
        range-wise all-reduce (efficiency 0.954; 22.34 ms with NCCL_MAX_CTAS=4), 21.37 ms with one 86.6 MB all-reduce
        after backward (0.972): exposing 0.6 ms of communication costs less than disturbing 10 ms of GEMMs.  The windowed
        mode keeps the collectives away from the persistent kernels instead."""
        super().__init__()
        self.module = module
        self.reducer = FlatGradReducer(process_group)
        if overlap is None:
            import os
            overlap = os.environ.get("SVIT_DDP_OVERLAP", DEFAULT_OVERLAP)
        self.overlap = {True: "range", False: "none", "1": "range", "0": "none", "": "none"}.get(overlap, overlap)
        if self.overlap not in ("none", "range", "window"):
            raise ValueError(f"overlap must be one of none / range / window (got {overlap!r})")
        self.min_window_numel = 1 << 16
        self._ready = None      # windowed mode: [lo, hi) of the flat buffer that is final but not yet all-reduced
        sit = getattr(module, "transformer", None)
        self._sit = module if hasattr(module, "stage_segment") else sit
        if self._sit is None or not hasattr(self._sit, "stage_segment"):
            raise TypeError("DataParallel wraps a B200 SiT or masked_patch_pretraining")
        self._sit._grad_hook = self._on_stage
        if self._sit is not module:
            module._grad_hook = self._on_small_buffer
        if broadcast_parameters and self.reducer.world > 1:
            dist.broadcast(self._sit._flat, src=0, group=process_group)
            self._sit.mark_weights_dirty()
            if self._sit is not module:
                dist.broadcast(module._flat, src=0, group=process_group)
                module.mark_weights_dirty()

    def _flush_ready(self, G):
        if self._ready is not None:
            lo, hi = self._ready
            self.reducer.reduce_range(G, lo, hi - lo)
            self._ready = None

    def _on_stage(self, sit, stage, G):
        if stage is None:                       # backward fully enqueued
            if self.overlap == "none":
                self.reducer.reduce_range(G, 0, G.numel())
            self._flush_ready(G)
            self.reducer.finish()
            return
        if stage == "all":                      # encoder-only backward: no stages reported
            if self.overlap != "none":
                self.reducer.reduce_range(G, 0, G.numel())
            return
        if self.overlap == "none":
            return
        if stage == sit.depth:
            self._ready = None                  # first callback of a backward: nothing can be left over from a failed one
        if stage >= STAGE_WINDOW:
            # (the head's 1.2 k gradients alone are not worth a collective: they ride with the last layer's range)
            if self.overlap == "window" and self._ready is not None and self._ready[1] - self._ready[0] >= self.min_window_numel:
                self._flush_ready(G)
            return
        start, numel = sit.stage_segment(stage)
        if self.overlap == "range":
            self.reducer.reduce_range(G, start, numel)
            return
        # windowed: the stages arrive from the end of the flat buffer towards its start (head, layers depth-1 .. 0, patch
        # embedding), so what is final and unsent stays ONE contiguous range
        if self._ready is None:
            self._ready = (start, start + numel)
        elif start + numel == self._ready[0]:
            self._ready = (start, self._ready[1])
        elif start == self._ready[1]:
            self._ready = (self._ready[0], start + numel)
        else:
            self._flush_ready(G)
            self._ready = (start, start + numel)

    def _on_small_buffer(self, module, MG):
        self.reducer.reduce_range(MG, 0, MG.numel())
        self.reducer.finish()

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
