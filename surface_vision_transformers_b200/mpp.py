"""Drop-in ``masked_patch_pretraining`` -- same constructor / forward contract as the reference's
``models/mpp.py`` (/root/reference/models/mpp.py:46-134): returns ``(mpp_loss, batch_out)``.

The masks are drawn on the host side with exactly the reference's torch RNG calls, in the reference's order
(mpp.py:25-43 and :85-111, mixed device + CPU generators), and handed to the fused sm_100a path, which applies the
corruption while packing the patch-embedding operand, runs the encoder, the decoder GEMM and the masked L2 loss.
"""
import math
import weakref

import torch
from torch import nn

from . import _lib
from ._lib import check, ptr, vp
from .sit import SiT, _stream

__all__ = ["masked_patch_pretraining", "get_mask_from_prob", "prob_mask_like", "draw_masks"]


def get_mask_from_prob(inputs, prob):
    # mpp.py:25-39 (inputs only provides batch, seq_len and device)
    batch, seq_len, device = inputs.shape[0], inputs.shape[1], inputs.device
    max_masked = math.ceil(prob * seq_len)
    rand = torch.rand((batch, seq_len), device=device)
    _, sampled_indices = rand.topk(max_masked, dim=-1)
    new_mask = torch.zeros((batch, seq_len), device=device)
    new_mask.scatter_(1, sampled_indices, 1)
    return new_mask.bool()


def prob_mask_like(inputs, prob):
    # mpp.py:41-43 -- CPU generator, as in the reference
    batch, seq_length = inputs.shape[0], inputs.shape[1]
    return torch.zeros((batch, seq_length)).float().uniform_(0, 1) < prob


class _Shape:
    """(B, N, K)-shaped stand-in so the mask helpers can be called without materialising 'b n (v c)'."""

    def __init__(self, b, n, k, device):
        self.shape = (b, n, k)
        self.device = device


def draw_masks(b, n, k, device, mask_prob, replace_prob, swap_prob):
    """RNG call order of mpp.py:85-111.  Returns (mask, swap_sel, swap_src, replace_sel)."""
    like = _Shape(b, n, k, device)
    mask = get_mask_from_prob(like, mask_prob)                                           # :85
    swap_sel = swap_src = None
    if swap_prob > 0:
        p = swap_prob / (1 - replace_prob)                                                 # :91
        random_patch_prob = prob_mask_like(like, p).to(device)                           # :94
        swap_sel = mask * (random_patch_prob == True)                                    # :97  # noqa: E712
        swap_src = torch.randint(0, n, (b, n), device=device)                            # :99
    tokens_to_mask = prob_mask_like(like, replace_prob).to(device)                       # :109
    replace_sel = (mask * tokens_to_mask) == True                                        # :111  # noqa: E712
    return mask, swap_sel, swap_src, replace_sel


def _u8(t):
    return None if t is None else t.to(torch.uint8).contiguous()


class _MPPFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, batch, masks, training, to_w, to_b, mask_token, *params):
        model = module.transformer
        B, dev = batch.shape[0], batch.device
        model._refresh_shadow(dev, force=bool(training))
        module._refresh_shadow(dev, force=bool(training))
        ctx.shadow_gen = model._shadow_gen
        lib = _lib.load()
        mask, swap_sel, swap_src, replace_sel = masks
        mask8, swap8, repl8 = _u8(mask), _u8(swap_sel), _u8(replace_sel)
        src64 = None if swap_src is None else swap_src.to(torch.int64).contiguous()
        nbytes = lib.svit_workspace_bytes(model._engine, B, 1 if training else 0, 1)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        T, K = model.num_patches + 1, model.num_channels * model.num_vertices
        out_full = torch.empty(B, T, K, dtype=torch.float32, device=dev)
        loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
        ctx.drop = model._next_dropout_state()   # emb_dropout (mpp.py:125) and the encoder's dropouts (mpp.py:128)
        model._apply_dropout_state(ctx.drop)
        check(lib.svit_mpp_forward(model._engine, ptr(model._flat), ptr(model._shadow), ptr(module._shadow),
                                   ptr(to_b), ptr(mask_token), ptr(ws), nbytes, ptr(batch), B, ptr(mask8), ptr(swap8),
                                   ptr(src64), ptr(repl8), ptr(loss_sum), ptr(out_full), 1 if training else 0,
                                   _stream(dev)), "svit_mpp_forward")
        count = mask8.sum().to(torch.float32) * K     # masked rows * K elements (mse_loss 'mean', mpp.py:132)
        loss = loss_sum / count
        if training:
            ctx.module = module
            ctx.ws = ws
            ctx.B = B
            ctx.dev = dev
            ctx.save_for_backward(batch, out_full, mask8, repl8 if repl8 is not None else mask8, count)
            ctx.has_replace = repl8 is not None
        batch_out = out_full[:, 1:, :]
        ctx.mark_non_differentiable(batch_out)
        return loss, batch_out

    @staticmethod
    def backward(ctx, dloss, _dout):
        module = ctx.module
        model = module.transformer
        lib = _lib.load()
        batch, out_full, mask8, repl8, count = ctx.saved_tensors
        model._check_shadow_gen(ctx.shadow_gen)
        coef = (dloss.float() * 2.0 / count).reshape(()).contiguous()
        G = torch.zeros_like(model._flat)
        MG = torch.zeros_like(module._flat)
        with torch.cuda.device(ctx.dev):
            hook = model._make_progress_hook(G)
            model._apply_dropout_state(ctx.drop)
            check(lib.svit_mpp_backward(model._engine, ptr(model._flat), ptr(model._shadow), ptr(module._shadow),
                                        ptr(ctx.ws), ctx.B, ptr(batch), ptr(out_full), ptr(mask8),
                                        ptr(repl8) if ctx.has_replace else vp(0), ptr(coef), ptr(G), ptr(MG), hook, vp(0),
                                        _stream(ctx.dev)), "svit_mpp_backward")
            model._finish_progress_hook(G)
            if module._grad_hook is not None:
                module._grad_hook(module, MG)
        ctx.ws = None
        K, D = module.dim_out, module.dim_in
        gw = MG[:K * D].view(K, D)
        gb = MG[K * D:K * D + K]
        gm = MG[K * D + K:K * D + 2 * K].view(1, 1, K)
        # SiT parameters that the MPP graph does not reach (mlp_head.*) get no gradient, like in the reference
        grads = list(model._grad_views(G))
        for i in range(len(grads) - 4, len(grads)):
            grads[i] = None
        return (None, None, None, None, gw, gb, gm) + tuple(grads)


class masked_patch_pretraining(nn.Module):
    # mpp.py:48-74
    def __init__(self, transformer, dim_in, dim_out, device, mask_prob=0.15, replace_prob=0.5, swap_prob=0.3,
                 channels=4, num_vertices=561):
        super().__init__()
        if not isinstance(transformer, SiT):
            raise TypeError("masked_patch_pretraining (B200) wraps the B200 SiT")
        self.transformer = transformer
        self.dim_out = dim_out
        self.dim_in = dim_in
        self.to_original = nn.Linear(dim_in, dim_out)
        self.to_original.to(device)
        self.mask_prob = mask_prob
        self.replace_prob = replace_prob
        self.swap_prob = swap_prob
        self.mask_token = nn.Parameter(torch.randn(1, 1, channels * num_vertices))
        if dim_out != transformer.num_channels * transformer.num_vertices or dim_in != transformer.dim or \
                channels * num_vertices != dim_out:
            raise ValueError("dim_in / dim_out / channels*num_vertices must match the wrapped SiT")
        self._flat = None
        self._shadow = None
        self._shadow_key = None
        self._grad_hook = None
        self._flatten()

    def _flatten(self):
        K, D = self.dim_out, self.dim_in
        ps = [self.to_original.weight, self.to_original.bias, self.mask_token]
        devs = {p.device for p in ps}
        dev = self.to_original.weight.device if len(devs) > 1 else devs.pop()
        flat = torch.zeros(K * D + 2 * K, dtype=torch.float32, device=dev)
        off = 0
        self._offsets = []
        for i, p in enumerate(ps):
            n = p.numel()
            v = flat[off:off + n].view(p.shape)
            v.copy_(p.data.to(device=dev, dtype=torch.float32))
            p.data = v
            # the fused optimizers (optim.py) update parameters that name their flat owner with ONE kernel launch
            p._svit_owner = weakref.ref(self)
            p._svit_index = i
            self._offsets.append((off, n))
            off += n
        self._plist = ps
        self._flat = flat
        self._shadow = None
        self._shadow_key = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if self._flat is not None:
            self._flatten()
        return out

    def mark_weights_dirty(self):
        self._shadow_key = None

    def load_state_dict(self, state_dict, strict=True, **kwargs):
        out = super().load_state_dict(state_dict, strict=strict, **kwargs)
        self.mark_weights_dirty()
        self.transformer.mark_weights_dirty()
        return out

    def _refresh_shadow(self, dev, force=False):
        if self._flat.device != dev:
            raise RuntimeError(f"masked_patch_pretraining parameters are on {self._flat.device}, input on {dev}; "
                               "call ssl.to(device) as tools/pretrain.py:258 does")
        key = (self._flat.data_ptr(), self._flat._version, self.to_original.weight._version)
        if self._shadow is not None and key == self._shadow_key and not force:
            return
        lib = _lib.load()
        eng = self.transformer._engine
        if self._shadow is None or self._shadow.device != dev:
            self._shadow = torch.empty(lib.svit_mpp_shadow_bytes(eng), dtype=torch.uint8, device=dev)
        check(lib.svit_mpp_prepare_weights(eng, ptr(self.to_original.weight), ptr(self._shadow), _stream(dev)),
              "svit_mpp_prepare_weights")
        self._shadow_key = key

    # mpp.py:77-134
    def forward(self, batch, masks=None, **kwargs):
        t = self.transformer
        batch = t._check_input(batch)
        b, n = batch.shape[0], t.num_patches
        if masks is None:
            masks = draw_masks(b, n, self.dim_out, batch.device, self.mask_prob, self.replace_prob, self.swap_prob)
        training = torch.is_grad_enabled()
        with torch.cuda.device(batch.device):   # the engine launches on the CURRENT device's stream
            loss, out = _MPPFunction.apply(self, batch, masks, training, self.to_original.weight, self.to_original.bias,
                                           self.mask_token, *t._plist)
        return loss, out
