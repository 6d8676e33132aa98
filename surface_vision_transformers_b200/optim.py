"""Fused optimizers over the flat parameter buffer (torch.optim.AdamW/Adam/SGD as used at
/root/reference/tools/train.py:228-241,291).

They are ordinary ``torch.optim.Optimizer`` subclasses (zero_grad / step / param_groups / state_dict work as the
reference's drivers expect).  Parameters that belong to a B200 ``SiT`` / ``masked_patch_pretraining`` are updated by
ONE kernel launch per module through the C ABI; per-parameter state tensors are views of flat moment buffers, so
``optimizer.state_dict()`` keeps torch's layout.  Parameters whose ``.grad`` is None are skipped exactly like torch
does (no moment update, no step increment).
"""
import ctypes

import torch
from torch.optim import Optimizer

from . import _lib
from ._lib import AdamSegment, ADAM_BLOCK_ELEMS, check, ptr, vp


def _owner_of(p):
    ref = getattr(p, "_svit_owner", None)
    return ref() if ref is not None else None


def _flat_grad_for(owner, params):
    """Returns a flat gradient tensor laid out like owner._flat.  Fast path: every .grad already is the canonical
    view of one flat buffer (what the fused backward produces); otherwise gradients are gathered into a new one."""
    first = next((p for p in params if p.grad is not None), None)
    if first is None:
        return None, [False] * len(params)
    active = [p.grad is not None for p in params]
    base_ptr = first.grad.data_ptr() - owner._offsets[first._svit_index][0] * 4
    ok = all((not a) or (p.grad.dtype == torch.float32 and p.grad.is_contiguous() and
                         p.grad.data_ptr() == base_ptr + owner._offsets[p._svit_index][0] * 4)
             for p, a in zip(params, active))
    if ok:
        storage = first.grad.untyped_storage()
        start = (base_ptr - storage.data_ptr()) // 4
        if start >= 0 and storage.nbytes() >= (start + owner._flat.numel()) * 4:
            G = torch.empty(0, dtype=torch.float32, device=first.grad.device)
            G.set_(storage, start, (owner._flat.numel(),))
            return G, active
    G = torch.zeros_like(owner._flat)
    for p, a in zip(params, active):
        if a:
            off, n = owner._offsets[p._svit_index]
            G[off:off + n].copy_(p.grad.reshape(-1))
    return G, active


def _carry_over(old, new):
    """Copies the moment buffers / step counters of a flat state whose parameter buffer was re-allocated."""
    if old is None or len(old.bufs) != len(new.bufs) or old.bufs[0].numel() != new.bufs[0].numel():
        return False
    for a, b in zip(old.bufs, new.bufs):
        b.copy_(a)
    new.steps = list(old.steps)
    return True


class _FlatState:
    """Flat moment buffers + host-side segment table of one SiT owned by an optimizer."""

    def __init__(self, owner, nbuf):
        self.flat_ptr = owner._flat.data_ptr()
        self.bufs = [torch.zeros_like(owner._flat) for _ in range(nbuf)]
        n = len(owner._plist)
        self.steps = [0] * n
        blocks = []
        for i, (off, numel) in enumerate(owner._offsets):
            for c in range((numel + ADAM_BLOCK_ELEMS - 1) // ADAM_BLOCK_ELEMS):
                blocks += [i, c]
        self.nblocks = len(blocks) // 2
        self.block_map = torch.tensor(blocks, dtype=torch.int32, device=owner._flat.device)
        self.seg_host = (AdamSegment * n)()
        self.seg_pinned = torch.empty(ctypes.sizeof(self.seg_host), dtype=torch.uint8).pin_memory() \
            if owner._flat.is_cuda else None
        self.seg_dev = torch.empty(ctypes.sizeof(self.seg_host), dtype=torch.uint8, device=owner._flat.device)
        self.table_key = None      # active pattern the device table was uploaded for (None: upload needed)
        self.has_momentum = False  # FusedSGD: the momentum buffer holds a value (a step ran or a state_dict was loaded)


class FusedAdamW(Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay); ``decoupled=False`` gives torch.optim.Adam (L2)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, decoupled=True,
                 grad_scale=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled)
        super().__init__(params, defaults)
        self.grad_scale = grad_scale
        self._flat_states = {}

    def _state_for(self, owner, rebind=True):
        st = self._flat_states.get(id(owner))
        if st is None or st.flat_ptr != owner._flat.data_ptr():
            old, st = st, _FlatState(owner, 2)
            _carry_over(old, st)     # the flat buffer was re-created (model.to(), .float()): keep the moments and steps
            self._flat_states[id(owner)] = st
            if not rebind:
                return st
            for p in owner._plist:
                off, n = owner._offsets[p._svit_index]
                self.state[p] = {"step": torch.tensor(0.0), "exp_avg": st.bufs[0][off:off + n].view(p.shape),
                                 "exp_avg_sq": st.bufs[1][off:off + n].view(p.shape)}
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            by_owner, loose = {}, []
            for p in group["params"]:
                o = _owner_of(p)
                if o is not None and hasattr(o, "_offsets") and hasattr(p, "_svit_index"):
                    by_owner.setdefault(id(o), (o, []))[1].append(p)
                else:
                    loose.append(p)
            for owner, ps in by_owner.values():
                if len(ps) != len(owner._plist):
                    loose += ps        # partial parameter sets take the generic path
                    continue
                ps = owner._plist
                G, active = _flat_grad_for(owner, ps)
                if G is None:
                    continue
                st = self._state_for(owner)
                dev = owner._flat.device
                stream = vp(torch.cuda.current_stream(dev).cuda_stream)
                # The step counters and bias corrections live in the DEVICE table and are advanced by a stream-ordered
                # kernel; the host uploads the table only when the active pattern changes (first step, a parameter
                # gaining / losing its gradient, load_state_dict).  Nothing step-dependent is staged through pinned
                # memory any more -- a host that runs several steps ahead of the GPU would overwrite it before the copy
                # executes -- and a captured CUDA graph of the step replays correctly (graphs.GraphedTrainStep).
                key = tuple(active)
                if st.table_key != key:
                    if torch.cuda.is_current_stream_capturing():
                        raise RuntimeError("FusedAdamW: the set of parameters with gradients changed during CUDA graph "
                                           "capture; run the warm-up steps with the same graph first")
                    for i, a in enumerate(active):
                        off, n = owner._offsets[i]
                        s = st.seg_host[i]
                        s.offset, s.numel, s.active, s.step = off, n, 1 if a else 0, st.steps[i]
                        s.bias_corr1 = s.bias_corr2 = 1.0
                    ctypes.memmove(st.seg_pinned.data_ptr(), ctypes.addressof(st.seg_host), ctypes.sizeof(st.seg_host))
                    st.seg_dev.copy_(st.seg_pinned)      # synchronous on purpose (rare)
                    st.table_key = key
                self._note_step(owner, st, active)
                with torch.cuda.device(dev):   # kernels launch on the current device
                    check(lib.svit_adamw_advance(ptr(st.seg_dev), len(ps), beta1, beta2, stream), "svit_adamw_advance")
                    check(lib.svit_adamw_step(ptr(owner._flat), ptr(G), ptr(st.bufs[0]), ptr(st.bufs[1]), ptr(st.seg_dev),
                                              len(ps), ptr(st.block_map), st.nblocks, group["lr"], beta1, beta2, group["eps"],
                                              group["weight_decay"], 1 if group["decoupled"] else 0, self.grad_scale,
                                              stream), "svit_adamw_step")
                owner.mark_weights_dirty()
            self._generic_adam(group, loose)
        return loss

    def _note_step(self, owner, st, active):
        """Host mirror of the device-side step counters (what ``state_dict()`` reports as ``state[p]['step']``)."""
        st.last_active = list(active)
        if torch.cuda.is_current_stream_capturing():
            return      # a capture records the step without executing it: the device counters do not move either
        for i, (p, a) in enumerate(zip(owner._plist, active)):
            if a:
                st.steps[i] += 1
                self.state[p]["step"] = torch.tensor(float(st.steps[i]))
        st.last_active = list(active)

    def note_graph_replay(self):
        """A captured training step was replayed: the device advanced its counters, mirror that on the host."""
        for st, owner in ((s_, o_) for s_, o_ in self._owners()):
            if getattr(st, "last_active", None) is not None:
                self._note_step(owner, st, st.last_active)

    def _owners(self):
        seen = {}
        for group in self.param_groups:
            for p in group["params"]:
                o = _owner_of(p)
                if o is not None and id(o) in self._flat_states:
                    seen[id(o)] = (self._flat_states[id(o)], o)
        return list(seen.values())

    def load_state_dict(self, state_dict):
        """torch's loader replaces the per-parameter state tensors; copy them back into the flat moment buffers (the
        fused kernel updates those), restore the step counters and force a re-upload of the device table."""
        super().load_state_dict(state_dict)
        for group in self.param_groups:
            owners = {}
            for p in group["params"]:
                o = _owner_of(p)
                if o is not None and hasattr(o, "_offsets") and hasattr(p, "_svit_index"):
                    owners[id(o)] = o
            for owner in owners.values():
                st = self._state_for(owner, rebind=False)
                for i, p in enumerate(owner._plist):
                    loaded = self.state.get(p, {})
                    off, n = owner._offsets[i]
                    if "exp_avg" in loaded:
                        st.bufs[0][off:off + n].copy_(loaded["exp_avg"].reshape(-1))
                        st.bufs[1][off:off + n].copy_(loaded["exp_avg_sq"].reshape(-1))
                        st.steps[i] = int(float(loaded.get("step", 0)))
                    self.state[p] = {"step": torch.tensor(float(st.steps[i])),
                                     "exp_avg": st.bufs[0][off:off + n].view(p.shape),
                                     "exp_avg_sq": st.bufs[1][off:off + n].view(p.shape)}
                st.table_key = None

    def _generic_adam(self, group, params):
        beta1, beta2 = group["betas"]
        for p in params:
            if p.grad is None:
                continue
            g = p.grad * self.grad_scale
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p)
                st["exp_avg_sq"] = torch.zeros_like(p)
            st["step"] += 1
            k = float(st["step"])
            if group["decoupled"]:
                p.mul_(1 - group["lr"] * group["weight_decay"])
            else:
                g = g.add(p, alpha=group["weight_decay"])
            st["exp_avg"].mul_(beta1).add_(g, alpha=1 - beta1)
            st["exp_avg_sq"].mul_(beta2).addcmul_(g, g, value=1 - beta2)
            denom = (st["exp_avg_sq"].sqrt() / (1 - beta2 ** k) ** 0.5).add_(group["eps"])
            p.addcdiv_(st["exp_avg"], denom, value=-group["lr"] / (1 - beta1 ** k))
            o = _owner_of(p)
            if o is not None:
                o.mark_weights_dirty()


class FusedSGD(Optimizer):
    """torch.optim.SGD semantics (momentum / dampening / nesterov / L2 weight decay) -- the reference's YAML default
    optimiser (config/SiT/training/hparams.yml:50-57)."""

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, grad_scale=1.0):
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(params, defaults)
        self.grad_scale = grad_scale
        self._flat_states = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            by_owner, loose = {}, []
            for p in group["params"]:
                o = _owner_of(p)
                if o is not None and hasattr(o, "_offsets") and hasattr(p, "_svit_index"):
                    by_owner.setdefault(id(o), (o, []))[1].append(p)
                else:
                    loose.append(p)
            for owner, ps in by_owner.values():
                full = len(ps) == len(owner._plist) and all(p.grad is not None for p in ps)
                if not full:
                    loose += ps
                    continue
                G, _ = _flat_grad_for(owner, owner._plist)
                st = self._flat_states.get(id(owner))
                if st is None or st.flat_ptr != owner._flat.data_ptr():
                    old, st = st, _FlatState(owner, 1)
                    st.has_momentum = _carry_over(old, st) and old.has_momentum   # re-flattened model: keep the momentum
                    self._flat_states[id(owner)] = st
                    self._bind_views(owner, st)
                first = not st.has_momentum   # torch: the first step initialises the buffer with the gradient itself
                st.has_momentum = True
                dev = owner._flat.device
                with torch.cuda.device(dev):   # kernels launch on the current device
                    check(lib.svit_sgd_step(ptr(owner._flat), ptr(G), ptr(st.bufs[0]), owner._flat.numel(), group["lr"],
                                            group["momentum"], group["dampening"], group["weight_decay"],
                                            1 if group["nesterov"] else 0, 1 if first else 0, self.grad_scale,
                                            vp(torch.cuda.current_stream(dev).cuda_stream)), "svit_sgd_step")
                owner.mark_weights_dirty()
            self._generic_sgd(group, loose)
        return loss

    def _bind_views(self, owner, st):
        for p in owner._plist:
            off, n = owner._offsets[p._svit_index]
            self.state[p] = {"momentum_buffer": st.bufs[0][off:off + n].view(p.shape)}

    def load_state_dict(self, state_dict):
        """torch's loader replaces ``state[p]['momentum_buffer']`` with fresh tensors; copy them back into the flat
        momentum buffer (the fused kernel updates that one), re-bind the views and keep the momentum (no 'first step')."""
        super().load_state_dict(state_dict)
        for group in self.param_groups:
            owners = {}
            for p in group["params"]:
                o = _owner_of(p)
                if o is not None and hasattr(o, "_offsets") and hasattr(p, "_svit_index"):
                    owners[id(o)] = o
            for owner in owners.values():
                loaded = [self.state.get(p, {}).get("momentum_buffer") for p in owner._plist]
                if not any(b is not None for b in loaded):
                    continue
                st = self._flat_states.get(id(owner))
                if st is None or st.flat_ptr != owner._flat.data_ptr():
                    st = _FlatState(owner, 1)
                    self._flat_states[id(owner)] = st
                for (off, n), b in zip(owner._offsets, loaded):
                    if b is not None:
                        st.bufs[0][off:off + n].copy_(b.reshape(-1))
                st.has_momentum = True
                self._bind_views(owner, st)

    def _generic_sgd(self, group, loose):
        """Parameters that do not belong to a flat B200 module (or a partial set of one): plain torch arithmetic."""
        for p in loose:
            if p.grad is None:
                continue
            g = (p.grad * self.grad_scale).add(p, alpha=group["weight_decay"])
            if group["momentum"] != 0:
                stp = self.state[p]
                if "momentum_buffer" not in stp:
                    stp["momentum_buffer"] = g.clone()
                else:
                    stp["momentum_buffer"].mul_(group["momentum"]).add_(g, alpha=1 - group["dampening"])
                g = g.add(stp["momentum_buffer"], alpha=group["momentum"]) if group["nesterov"] else stp["momentum_buffer"]
            p.add_(g, alpha=-group["lr"])
            o = _owner_of(p)
            if o is not None:
                o.mark_weights_dirty()
