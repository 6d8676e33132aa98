"""CUDA-graph capture of the SiT step ("CUDA-graph capture of the step").

The engine behind ``SiT`` only enqueues kernels on the current stream, takes its workspace from the caching allocator
and keeps every step-dependent quantity on the device (the AdamW step counters are advanced by a kernel), so a whole
inference pass -- or forward + loss + backward + fused optimizer step -- can be captured once and replayed with one
launch.  At the benchmark batch (256 per GPU) the step is GPU-bound and a graph changes nothing; at small batches the
~230 launches of a step cost more host time than the kernels take (SiT-tiny, batch 16: 2.8 -> 2.3 ms per training step,
0.83 -> 0.64 ms per inference pass on a B200).

Limits: single process (no ``DataParallel``), dropout 0 (the mask offset is host state), a constant learning rate between
captures (it is a kernel argument; call ``recapture()`` after a scheduler step), fixed input shapes.
"""
import torch

from .ddp import DataParallel
from .sit import SiT

__all__ = ["GraphedInference", "GraphedTrainStep"]


def _check_model(model):
    if isinstance(model, DataParallel):
        raise TypeError("CUDA-graph capture is single-process: pass the SiT, not the DataParallel wrapper")
    core = model if isinstance(model, SiT) else getattr(model, "transformer", None)
    if not isinstance(core, SiT):
        raise TypeError("expected a B200 SiT or masked_patch_pretraining")
    if core._drop_p > 0.0 or core._emb_drop_p > 0.0:
        raise NotImplementedError("CUDA-graph capture with dropout > 0: the mask offset is host state")
    return core


class GraphedInference:
    """``y = GraphedInference(model, example_input)(x)``: ``model.eval()`` forward replayed from a CUDA graph.
    The returned tensor is a static buffer that the next call overwrites."""

    def __init__(self, model, example_input, warmup=3):
        _check_model(model)
        self.model = model
        self.x = example_input.detach().clone().contiguous().float()
        model.eval()
        side = torch.cuda.Stream(self.x.device)
        side.wait_stream(torch.cuda.current_stream(self.x.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                model(self.x)
        torch.cuda.current_stream(self.x.device).wait_stream(side)
        # capture the bf16 weight-shadow refresh with the pass: a replay then always reads the CURRENT fp32 weights (a graph
        # captured over clean shadows would keep using them after the weights were trained or reloaded)
        core = model if isinstance(model, SiT) else model.transformer
        core.mark_weights_dirty()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.y = model(self.x)
        core.mark_weights_dirty()

    def __call__(self, x):
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y


class GraphedTrainStep:
    """``loss = GraphedTrainStep(model, optimizer, criterion, x, y)(x, y)``: zero_grad + forward + criterion + backward +
    ``optimizer.step()`` replayed from one CUDA graph.  ``optimizer`` must be a ``FusedAdamW`` / ``FusedSGD`` over the
    model's parameters; ``criterion(outputs, targets)`` any capturable torch expression.  Returns the (static) loss
    tensor of the replayed step."""

    def __init__(self, model, optimizer, criterion, example_input, example_target, warmup=3):
        _check_model(model)
        self.model, self.optimizer, self.criterion = model, optimizer, criterion
        self.x = example_input.detach().clone().contiguous().float()
        self.t = example_target.detach().clone()
        self.warmup = warmup
        self.recapture()

    def _step(self):
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(self.x)
        loss = self.criterion(out, self.t)
        loss.backward()
        self.optimizer.step()
        return loss

    def recapture(self):
        """(Re)captures the step, e.g. after the learning rate changed.  The warm-up iterations are REAL optimizer
        steps on the example batch (they also bring the optimizer's device tables into their steady state); the capture
        itself only records -- it does not execute a step."""
        dev = self.x.device
        self.model.train()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        self.lr = [g["lr"] for g in self.optimizer.param_groups]
        return self

    def __call__(self, x, t):
        if [g["lr"] for g in self.optimizer.param_groups] != self.lr:
            raise RuntimeError("the learning rate changed since capture: call recapture() first")
        self.x.copy_(x, non_blocking=True)
        self.t.copy_(t, non_blocking=True)
        self.graph.replay()
        note = getattr(self.optimizer, "note_graph_replay", None)
        if note is not None:
            note()
        return self.loss
