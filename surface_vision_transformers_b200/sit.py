"""Drop-in ``SiT`` -- same constructor, attributes, ``forward`` and ``state_dict`` as the reference's
``models/sit.py::SiT`` (/root/reference/models/sit.py:25-82), with forward and backward executed by the
hand-written sm_100a kernels behind the C ABI in ``include/svit_b200.h``.

Host code is plumbing only: parameters live in ONE flat fp32 buffer (views keep the reference's names and
shapes, so checkpoints round-trip), activations live in a per-call workspace, and the whole forward (or
backward) is a single C call on the current CUDA stream.  There is no CPU / eager fallback: tensors that are
not on a CUDA device raise.
"""
import ctypes
import weakref

import torch
from torch import nn

from . import _lib
from ._lib import SvitConfig, check, ptr, vp

__all__ = ["SiT", "Transformer"]


class Rearrange(nn.Module):
    """'b c n v -> b n (v c)' (models/sit.py:49); parameter-free placeholder at index 0 of to_patch_embedding."""

    def forward(self, x):
        b, c, n, v = x.shape
        return x.permute(0, 2, 3, 1).reshape(b, n, v * c)


class _Holder(nn.Module):
    """Parameter container mirroring a vit_pytorch sub-module; the arithmetic runs in the fused engine."""

    def forward(self, *a, **k):
        raise RuntimeError("this sub-module only holds parameters; call SiT(...) or SiT.transformer(x) instead")


class Attention(_Holder):
    def __init__(self, dim, heads, dim_head, dropout):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))


class FeedForward(_Holder):
    def __init__(self, dim, hidden_dim, dropout):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class PreNorm(_Holder):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class Transformer(nn.Module):
    """Same parameter layout as ``vit_pytorch.vit.Transformer`` (keys pinned by utils/utils.py:18-33).
    ``forward(x)`` runs the fused encoder: x (B,T,D) fp32 -> (B,T,D) fp32."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                PreNorm(dim, Attention(dim, heads, dim_head, dropout)),
                PreNorm(dim, FeedForward(dim, mlp_dim, dropout)),
            ]))
        self._owner = None

    def forward(self, x, **kwargs):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise RuntimeError("Transformer is not attached to a SiT")
        return owner._encoder(x)


def _stream(device):
    return vp(torch.cuda.current_stream(device).cuda_stream)


class _SiTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, img, mesh_args, *params):
        # mesh_args: None for pre-patched input (B,C,N,V), or (table, n_mesh, ch_mean, ch_std) for raw meshes (B,C,n_mesh):
        # the gather (tools/preprocessing.py:79-84) and z-score (:72) then run inside the patch-packing kernel
        B = img.shape[0]
        dev = img.device
        model._refresh_shadow(dev, force=True)
        ctx.shadow_gen = model._shadow_gen
        lib = _lib.load()
        nbytes = lib.svit_workspace_bytes(model._engine, B, 1, 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty(B, model.num_classes, dtype=torch.float32, device=dev)
        ctx.drop = model._next_dropout_state()
        model._apply_dropout_state(ctx.drop)
        table, n_mesh, ch_mean, ch_std = mesh_args if mesh_args is not None else (None, 0, None, None)
        check(lib.svit_forward_ex(model._engine, ptr(model._flat), ptr(model._shadow), ptr(ws), nbytes, ptr(img),
                                  1 if img.dtype == torch.bfloat16 else 0, B, ptr(table), n_mesh, ptr(ch_mean), ptr(ch_std),
                                  ptr(out), 1, _stream(dev)), "svit_forward")
        ctx.model = model
        ctx.ws = ws
        ctx.B = B
        ctx.dev = dev
        return out

    @staticmethod
    def backward(ctx, dout):
        model = ctx.model
        lib = _lib.load()
        model._check_shadow_gen(ctx.shadow_gen)
        dout = dout.contiguous().float()
        G = torch.zeros_like(model._flat)
        with torch.cuda.device(ctx.dev):
            hook = model._make_progress_hook(G)
            model._apply_dropout_state(ctx.drop)
            check(lib.svit_backward(model._engine, ptr(model._flat), ptr(model._shadow), ptr(ctx.ws), ctx.B, ptr(dout),
                                    ptr(G), hook, vp(0), _stream(ctx.dev)), "svit_backward")
            model._finish_progress_hook(G)
        ctx.ws = None
        return (None, None, None) + model._grad_views(G)


class _EncoderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *params):
        B = x.shape[0]
        dev = x.device
        model._refresh_shadow(dev, force=True)
        ctx.shadow_gen = model._shadow_gen
        lib = _lib.load()
        training = 1
        nbytes = lib.svit_workspace_bytes(model._engine, B, training, 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        y = torch.empty_like(x)
        ctx.drop = model._next_dropout_state(emb=False)
        model._apply_dropout_state(ctx.drop)
        check(lib.svit_encoder_forward(model._engine, ptr(model._flat), ptr(model._shadow), ptr(ws), nbytes, ptr(x), B,
                                       ptr(y), training, _stream(dev)), "svit_encoder_forward")
        ctx.model = model
        ctx.ws = ws
        ctx.B = B
        ctx.dev = dev
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        model = ctx.model
        lib = _lib.load()
        (x,) = ctx.saved_tensors
        model._check_shadow_gen(ctx.shadow_gen)
        dy = dy.contiguous().float()
        G = torch.zeros_like(model._flat)
        dx = torch.empty_like(x) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(ctx.dev):
            model._apply_dropout_state(ctx.drop)
            check(lib.svit_encoder_backward(model._engine, ptr(model._flat), ptr(model._shadow), ptr(ctx.ws), ctx.B,
                                            ptr(x), ptr(dy), ptr(dx), ptr(G), _stream(ctx.dev)), "svit_encoder_backward")
            if model._grad_hook is not None:
                # the encoder-only backward has no per-stage progress callback: reduce the whole buffer at the end
                model._grad_hook(model, "all", G)
                model._grad_hook(model, None, G)
        ctx.ws = None
        return (None, dx) + model._grad_views(G)


class SiT(nn.Module):
    # models/sit.py:26-39 -- keyword-only constructor, same defaults
    def __init__(self, *, dim, depth, heads, mlp_dim, pool='cls', num_patches=20, num_classes=1, num_channels=4,
                 num_vertices=2145, dim_head=64, dropout=0., emb_dropout=0.):
        super().__init__()
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'
        if not (0. <= dropout < 1.) or not (0. <= emb_dropout < 1.):
            raise ValueError(f"dropout probabilities must be in [0, 1) (got {dropout}, {emb_dropout})")
        if dim_head != 64:
            raise NotImplementedError("the fused attention kernel requires dim_head == 64")
        patch_dim = num_channels * num_vertices
        self.to_patch_embedding = nn.Sequential(Rearrange(), nn.Linear(patch_dim, dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))

        self.dim, self.depth, self.heads, self.mlp_dim = dim, depth, heads, mlp_dim
        self.num_patches, self.num_classes = num_patches, num_classes
        self.num_channels, self.num_vertices = num_channels, num_vertices
        self.transformer._owner = weakref.ref(self)
        # dropout > 0: counter-based masks generated inside the engine (include/svit_b200.h)
        self._drop_p, self._emb_drop_p = float(dropout), float(emb_dropout)
        self._drop_seed = None     # drawn from torch's seed at first use; see set_dropout_seed
        self._drop_step = 0

        cfg = SvitConfig(dim, depth, heads, dim_head, mlp_dim, num_patches, num_vertices, num_channels, num_classes,
                         1 if pool == 'mean' else 0)
        lib = _lib.load()
        eng = lib.svit_create(ctypes.byref(cfg))
        if not eng:
            raise _lib.SvitError("svit_create failed: " + lib.svit_last_error().decode())
        self._engine = vp(eng)
        self._finalizer = weakref.finalize(self, lib.svit_destroy, self._engine)
        self._flat = None
        self._shadow = None
        self._shadow_key = None
        self._shadow_dirty_gen = 0
        self._shadow_gen = 0     # bumped every time the bf16 shadows are re-derived (backward checks it, see below)
        self._grad_hook = None   # set by ddp.DataParallel: called as hook(stage, G) while backward is being enqueued
        self._hook_error = None
        self._flatten()

    # ------------------------------------------------------------------ parameters
    def _canonical_params(self):
        ps = [self.pos_embedding, self.cls_token, self.to_patch_embedding[1].weight, self.to_patch_embedding[1].bias]
        for attn, ff in self.transformer.layers:
            ps += [attn.norm.weight, attn.norm.bias, attn.fn.to_qkv.weight, attn.fn.to_out[0].weight,
                   attn.fn.to_out[0].bias, ff.norm.weight, ff.norm.bias, ff.fn.net[0].weight, ff.fn.net[0].bias,
                   ff.fn.net[3].weight, ff.fn.net[3].bias]
        ps += [self.mlp_head[0].weight, self.mlp_head[0].bias, self.mlp_head[1].weight, self.mlp_head[1].bias]
        return ps

    def _flatten(self):
        """(Re)creates the flat fp32 buffer on the parameters' device and re-points every parameter at its slice."""
        lib = _lib.load()
        ps = self._canonical_params()
        assert len(ps) == lib.svit_num_params(self._engine)
        dev = ps[0].device
        flat = torch.zeros(lib.svit_flat_numel(self._engine), dtype=torch.float32, device=dev)
        self._offsets = []
        for i, p in enumerate(ps):
            off, n = lib.svit_param_offset(self._engine, i), lib.svit_param_numel(self._engine, i)
            assert n == p.numel(), (i, n, p.shape)
            view = flat[off:off + n].view(p.shape)
            view.copy_(p.data.to(torch.float32))
            p.data = view
            p._svit_owner = weakref.ref(self)
            p._svit_index = i
            self._offsets.append((off, n))
        self._flat = flat
        self._plist = ps
        self._shadow = None
        self._shadow_key = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if self._flat is not None:
            self._flatten()
        return out

    def _weights_version(self):
        return (self._flat.data_ptr(), self._flat._version, sum(p._version for p in self._plist))

    def set_check_mode(self, on=True):
        """fp32 check mode (include/svit_b200.h): the same forward / backward with every operand in fp32 on the CUDA
        cores instead of bf16 on the tensor cores -- for verification at a 1e-4 tolerance, not for speed."""
        check(_lib.load().svit_set_check_mode(self._engine, 1 if on else 0), "svit_set_check_mode")
        return self

    def set_dropout_seed(self, seed, step=0):
        """Fixes the (seed, step) of the dropout masks: forward number k after this call uses offset step + k."""
        self._drop_seed, self._drop_step = int(seed) & 0xFFFFFFFFFFFFFFFF, int(step)
        return self

    def _next_dropout_state(self, emb=True):
        """(p, emb_p, seed, offset) of the forward about to run; like nn.Dropout, active only in .train() mode."""
        p, emb_p = (self._drop_p, self._emb_drop_p if emb else 0.) if self.training else (0., 0.)
        if p == 0. and emb_p == 0.:
            return (0., 0., 0, 0)
        if self._drop_seed is None:
            rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
            self._drop_seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * rank) & 0xFFFFFFFFFFFFFFFF
        state = (p, emb_p, self._drop_seed, self._drop_step)
        self._drop_step += 1
        return state

    def _apply_dropout_state(self, state):
        check(_lib.load().svit_set_dropout(self._engine, state[0], state[1], state[2], state[3]), "svit_set_dropout")

    @property
    def check_mode(self):
        return bool(_lib.load().svit_get_check_mode(self._engine))

    def mark_weights_dirty(self):
        """The fp32 masters changed behind autograd's back: the bf16 shadows are re-derived before the next forward.
        Called by the fused optimizers and by ``load_state_dict``.  Every grad-enabled forward re-derives them anyway
        (in a training loop the optimizer has just changed the weights, so that costs nothing extra); only INFERENCE
        forwards trust the version counters, which writes through ``p.data`` (EMA / SWA swaps, ``p.data.copy_``) do
        not bump -- call this method after such writes."""
        self._shadow_key = None

    def load_state_dict(self, state_dict, strict=True, **kwargs):
        out = super().load_state_dict(state_dict, strict=strict, **kwargs)
        self.mark_weights_dirty()
        return out

    def _check_shadow_gen(self, gen):
        """Backward reads the same shadow buffer the forward used; if another forward re-derived it from CHANGED
        weights in between, the gradients would silently belong to the new weights -- refuse instead."""
        if gen != self._shadow_gen and self._shadow_changed_since(gen):
            raise RuntimeError("SiT: the weights were modified between this forward and its backward (the bf16 weight "
                               "shadows were re-derived from different values); run backward before updating the weights")

    def _shadow_changed_since(self, gen):
        return self._shadow_dirty_gen > gen

    def _refresh_shadow(self, dev, force=False):
        if self._flat.device != dev or dev.type != 'cuda':
            raise RuntimeError(f"SiT parameters are on {self._flat.device} but the input is on {dev}: the sm_100a path "
                               "needs both on the same CUDA device (there is no CPU fallback)")
        if self._flat.dtype != torch.float32:
            raise RuntimeError("SiT parameters must be fp32 master weights")
        key = self._weights_version()
        clean = self._shadow is not None and key == self._shadow_key
        if clean and not force:
            return
        lib = _lib.load()
        if self._shadow is None or self._shadow.device != dev:
            self._shadow = torch.empty(lib.svit_shadow_bytes(self._engine), dtype=torch.uint8, device=dev)
        check(lib.svit_prepare_weights(self._engine, ptr(self._flat), ptr(self._shadow), _stream(dev)),
              "svit_prepare_weights")
        self._shadow_key = key
        self._shadow_gen += 1
        if not clean:
            self._shadow_dirty_gen = self._shadow_gen   # last generation derived from weights known to have changed

    def _grad_views(self, G):
        return tuple(G[off:off + n].view(p.shape) for (off, n), p in zip(self._offsets, self._plist))

    # ------------------------------------------------------------------ DDP progress hook plumbing
    def _make_progress_hook(self, G):
        if self._grad_hook is None:
            return vp(0)
        hook = self._grad_hook

        def cb(stage, _user):
            # ctypes prints and then SWALLOWS exceptions raised inside a callback: stash the first one, it is re-raised
            # by _finish_progress_hook once svit_backward has returned (a failed all-reduce must not pass silently)
            if self._hook_error is not None:
                return
            try:
                hook(self, stage, G)
            except BaseException as e:   # noqa: BLE001
                self._hook_error = e

        self._hook_error = None
        self._cb_keepalive = _lib.PROGRESS_FN(cb)
        return ctypes.cast(self._cb_keepalive, vp)

    def _finish_progress_hook(self, G):
        self._cb_keepalive = None
        err, self._hook_error = self._hook_error, None
        if err is not None:
            raise RuntimeError("gradient hook failed during backward (replicas would diverge)") from err
        if self._grad_hook is not None:
            self._grad_hook(self, None, G)   # stage None == "backward fully enqueued": wait for outstanding work

    def stage_segment(self, stage):
        """(offset, numel) of the flat-buffer range whose gradients are final after `stage`
        (depth = head, l = encoder layer l, -1 = patch embedding / pos / cls)."""
        if stage == self.depth:
            a = self._offsets[4 + 11 * self.depth][0]
            return a, self._flat.numel() - a
        if stage >= 0:
            a = self._offsets[4 + 11 * stage][0]
            b = self._offsets[4 + 11 * (stage + 1)][0]
            return a, b - a
        return 0, self._offsets[4][0]

    # ------------------------------------------------------------------ forward paths
    def _check_input(self, img):
        if img.dim() != 4 or img.shape[1] != self.num_channels or img.shape[2] != self.num_patches or \
                img.shape[3] != self.num_vertices:
            raise ValueError(f"expected input (B,{self.num_channels},{self.num_patches},{self.num_vertices}), "
                             f"got {tuple(img.shape)}")
        if not img.is_cuda:
            raise RuntimeError("SiT (B200) needs a CUDA input: there is no CPU fallback")
        # bf16 batches (staged as bf16 on the host: half the host-to-device bytes) are consumed as they are -- the patch
        # packing kernel rounds fp32 input to bf16 anyway, so the results are bit-identical; anything else becomes fp32
        if img.dtype == torch.bfloat16 and not _lib.load().svit_get_check_mode(self._engine):
            return img.contiguous()
        return img.contiguous().float()

    def forward(self, img):
        img = self._check_input(img)
        if img.shape[0] == 0:   # an empty batch (e.g. an empty last shard): the reference returns an empty (0, classes) tensor too
            return img.new_zeros((0, self.num_classes)) + 0.0 * self.mlp_head[1].bias.sum()
        with torch.cuda.device(img.device):   # the engine launches on the CURRENT device's stream
            if torch.is_grad_enabled() and any(p.requires_grad for p in self._plist):
                return _SiTFunction.apply(self, img, None, *self._plist)
            return self.infer(img)

    @torch.no_grad()
    def infer(self, img, table=None, n_mesh=0, ch_mean=None, ch_std=None):
        """Forward without autograd (eval / torch.no_grad()); small ping-pong workspace."""
        B, dev = img.shape[0], img.device
        self._refresh_shadow(dev)
        lib = _lib.load()
        nbytes = lib.svit_workspace_bytes(self._engine, B, 0, 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty(B, self.num_classes, dtype=torch.float32, device=dev)
        self._apply_dropout_state(self._next_dropout_state())   # .train() under no_grad still drops, as nn.Dropout does
        check(lib.svit_forward_ex(self._engine, ptr(self._flat), ptr(self._shadow), ptr(ws), nbytes, ptr(img),
                                  1 if img.dtype == torch.bfloat16 else 0, B, ptr(table), n_mesh, ptr(ch_mean), ptr(ch_std),
                                  ptr(out), 0, _stream(dev)), "svit_forward")
        return out

    def forward_mesh(self, mesh, table, ch_mean=None, ch_std=None):
        """raw ico-6 mesh (B,C,40962) + gather table (V,N) int32 -> prediction; the patch gather
        (tools/preprocessing.py:79-84) and optional z-score (:72) are fused into the patch packing kernel that writes
        the embedding GEMM's bf16 operand -- no gathered fp32 copy of the input exists.  Works with autograd (training
        straight from raw meshes: the backward pass reads the packed operand the forward left in its workspace)."""
        if mesh.dim() != 3 or mesh.shape[1] != self.num_channels:
            raise ValueError(f"expected raw meshes (B,{self.num_channels},n_vertices), got {tuple(mesh.shape)}")
        if not mesh.is_cuda:
            raise RuntimeError("SiT (B200) needs a CUDA input: there is no CPU fallback")
        if tuple(table.shape) != (self.num_vertices, self.num_patches) or table.dtype != torch.int32:
            raise ValueError(f"gather table must be int32 ({self.num_vertices},{self.num_patches}), got {table.dtype} "
                             f"{tuple(table.shape)}")
        mesh = mesh.contiguous().float()
        table = table.contiguous()
        if ch_mean is not None:
            ch_mean = torch.as_tensor(ch_mean, dtype=torch.float32, device=mesh.device).contiguous()
            ch_std = torch.as_tensor(ch_std, dtype=torch.float32, device=mesh.device).contiguous()
        with torch.cuda.device(mesh.device):
            if mesh.shape[0] > 0 and torch.is_grad_enabled() and any(p.requires_grad for p in self._plist):
                return _SiTFunction.apply(self, mesh, (table, mesh.shape[-1], ch_mean, ch_std), *self._plist)
            return self.infer(mesh, table=table, n_mesh=mesh.shape[-1], ch_mean=ch_mean, ch_std=ch_std)

    def _encoder(self, x):
        if not x.is_cuda:
            raise RuntimeError("SiT.transformer (B200) needs a CUDA input: there is no CPU fallback")
        x = x.contiguous().float()
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self._plist)):
            with torch.cuda.device(x.device):
                return _EncoderFunction.apply(self, x, *self._plist)
        with torch.no_grad(), torch.cuda.device(x.device):
            B, dev = x.shape[0], x.device
            self._refresh_shadow(dev)
            lib = _lib.load()
            nbytes = lib.svit_workspace_bytes(self._engine, B, 0, 0)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            y = torch.empty_like(x)
            self._apply_dropout_state(self._next_dropout_state(emb=False))
            check(lib.svit_encoder_forward(self._engine, ptr(self._flat), ptr(self._shadow), ptr(ws), nbytes, ptr(x), B,
                                           ptr(y), 0, _stream(dev)), "svit_encoder_forward")
            return y
