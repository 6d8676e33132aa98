"""Fused regression criterion: ``nn.MSELoss(reduction='mean')`` / ``nn.L1Loss()`` on
``outputs.squeeze()`` vs ``targets`` exactly as the reference's training loop applies it (tools/train.py:245-248, :288).

``regression_loss(outputs, targets, l1loss=False)`` is ONE kernel launch for the scalar loss AND d loss / d outputs;
its backward multiplies that stored gradient by the incoming scalar.  The reference's own spelling
(``criterion(outputs.squeeze(), targets)`` with torch's modules) keeps working on the drop-in model; this is the
launch-free-backward version the training loop of this package uses.
"""
import torch

from . import _lib
from ._lib import check, ptr, vp

__all__ = ["regression_loss", "RegressionLoss"]


class _RegressionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, targets, l1):
        out = outputs.reshape(-1).contiguous().float()
        tgt = targets.reshape(-1).contiguous().float()
        if out.numel() != tgt.numel():
            raise ValueError(f"outputs {tuple(outputs.shape)} and targets {tuple(targets.shape)} do not match")
        if not out.is_cuda or tgt.device != out.device:
            raise RuntimeError("regression_loss needs CUDA tensors on one device (no CPU fallback)")
        loss = torch.empty((), dtype=torch.float32, device=out.device)
        dout = torch.empty_like(out)
        with torch.cuda.device(out.device):
            check(_lib.load().svit_regression_loss(ptr(out), ptr(tgt), out.numel(), 1 if l1 else 0, ptr(loss), ptr(dout),
                                                   vp(torch.cuda.current_stream(out.device).cuda_stream)), "svit_regression_loss")
        ctx.save_for_backward(dout)
        ctx.shape = outputs.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dout,) = ctx.saved_tensors
        return (dout * dloss).view(ctx.shape), None, None


def regression_loss(outputs, targets, l1loss=False):
    """mean((outputs.squeeze() - targets)^2), or mean(|.|) with ``l1loss`` (tools/train.py:245-248, :288)."""
    return _RegressionLoss.apply(outputs, targets, bool(l1loss))


class RegressionLoss(torch.nn.Module):
    """Module form: ``criterion = RegressionLoss(l1loss)``; ``criterion(outputs.squeeze(), targets)``."""

    def __init__(self, l1loss=False):
        super().__init__()
        self.l1loss = bool(l1loss)

    def forward(self, outputs, targets):
        return regression_loss(outputs, targets, self.l1loss)
