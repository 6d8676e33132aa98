"""CPU tests of the host side: the C-ABI library loads and exports every symbol of include/svit_b200.h, the
parameter layout / state_dict contract matches the reference, host-side mask drawing reproduces the reference's
RNG order, and the data-parallel gradient reducer works across 2 gloo ranks.  No kernel is launched here."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from helpers import CASES, ROOT, seeded_state
from oracle import reference_loader
from oracle.sit_oracle import OracleMPP, OracleSiT

import surface_vision_transformers_b200 as svit
from surface_vision_transformers_b200 import _lib

needs_reference = pytest.mark.skipif(not reference_loader.available(), reason="/root/reference not present")


def header_functions():
    h = open(os.path.join(ROOT, "include", "svit_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    names = set(re.findall(r"\b(svit_[a-z_0-9]+)\s*\(", h))
    return names - {"svit_progress_fn"}


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib._build.LIB if os.path.exists(_lib._build.LIB) else _lib._build.build())
    names = header_functions()
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/svit_b200.h but not exported"
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    assert _lib.load().svit_version() >= 100


def test_library_is_sm100a_tcgen05():
    """The shipped binary contains sm_100a SASS with tcgen05 MMA / TMEM / TMA instructions."""
    lib = _lib._build.LIB
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in out.stdout, mnemonic


def test_engine_rejects_unsupported_configs():
    lib = _lib.load()
    bad = _lib.SvitConfig(384, 12, 6, 32, 1536, 320, 153, 4, 1, 0)      # dim_head != 64
    assert not lib.svit_create(ctypes.byref(bad))
    assert b"dim_head" in lib.svit_last_error()
    bad = _lib.SvitConfig(384, 12, 6, 64, 1536, 1280, 45, 4, 1, 0)      # T = 1281 > 384
    assert not lib.svit_create(ctypes.byref(bad))
    assert b"sequence length" in lib.svit_last_error()
    with pytest.raises(NotImplementedError):
        svit.SiT(dim=128, depth=1, heads=2, mlp_dim=128, dropout=0.1)
    with pytest.raises(AssertionError):
        svit.SiT(dim=128, depth=1, heads=2, mlp_dim=128, pool="max")


@pytest.mark.parametrize("preset", [dict(dim=192, heads=3, mlp_dim=768), dict(dim=384, heads=6, mlp_dim=1536),
                                    dict(dim=768, heads=12, mlp_dim=3072)])
@pytest.mark.parametrize("patching", [dict(num_patches=320, num_vertices=153), dict(num_patches=80, num_vertices=561)])
def test_parameter_layout_and_counts(preset, patching):
    """Parameter counts of BASELINE.md section 2 and the flat layout of include/svit_b200.h."""
    expected = {(192, 320): 5511553, (192, 80): 5778817, (384, 320): 21639937, (384, 80): 22174465,
                (768, 320): 85747201, (768, 80): 86816257}
    m = svit.SiT(depth=12, **preset, **patching)
    n = sum(p.numel() for p in m.parameters())
    assert n == expected[(preset["dim"], patching["num_patches"])]
    assert len(list(m.parameters())) == 140
    lib = _lib.load()
    assert lib.svit_num_params(m._engine) == 140
    prev_end = 0
    for i, p in enumerate(m._plist):
        off, num = lib.svit_param_offset(m._engine, i), lib.svit_param_numel(m._engine, i)
        assert num == p.numel() and off % 64 == 0 and off >= prev_end
        assert p.data_ptr() == m._flat.data_ptr() + 4 * off          # parameters are views of the flat buffer
        prev_end = off + num
    assert lib.svit_flat_numel(m._engine) >= prev_end
    a = lib.svit_workspace_bytes(m._engine, 8, 1, 0)
    b = lib.svit_workspace_bytes(m._engine, 16, 1, 0)
    c = lib.svit_workspace_bytes(m._engine, 16, 0, 0)
    assert 0 < c < a < b


def test_state_dict_contract_and_roundtrip(tmp_path):
    cfg = CASES["sit_cls"]["cfg"]
    ours = svit.SiT(**cfg)
    orc = OracleSiT(**cfg)
    assert list(ours.state_dict().keys()) == list(orc.state_dict().keys())
    for k, v in orc.state_dict().items():
        assert ours.state_dict()[k].shape == v.shape and ours.state_dict()[k].dtype == torch.float32, k
    sd = seeded_state(orc, 1)
    ours.load_state_dict(sd)
    for k, v in sd.items():
        assert torch.equal(ours.state_dict()[k], v)
    # still views of one flat buffer after loading (copy_ in place), and a checkpoint round-trips
    assert all(p.data_ptr() == ours._flat.data_ptr() + 4 * off for p, (off, _) in zip(ours._plist, ours._offsets))
    f = tmp_path / "checkpoint.pth"
    torch.save(ours.state_dict(), f)                                   # tools/train.py:361-363
    orc.load_state_dict(torch.load(f))                                 # tools/testing.py:68
    res = ours.load_state_dict(torch.load(f), strict=False)            # tools/train.py:216
    assert not res.missing_keys and not res.unexpected_keys


@needs_reference
def test_state_dict_keys_equal_reference():
    SiT, MPP, _ = reference_loader.load_reference_models()
    cfg = CASES["mpp"]["cfg"]
    ref = SiT(**cfg)
    ours = svit.SiT(**cfg)
    assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in ours.named_parameters()]
    K = cfg["num_channels"] * cfg["num_vertices"]
    kw = dict(dim_in=cfg["dim"], dim_out=K, device="cpu", channels=cfg["num_channels"], num_vertices=cfg["num_vertices"])
    r2 = MPP(transformer=ref, **kw)
    o2 = svit.masked_patch_pretraining(transformer=ours, **kw)
    assert list(r2.state_dict().keys()) == list(o2.state_dict().keys())
    for k, v in r2.state_dict().items():
        assert o2.state_dict()[k].shape == v.shape, k
    # attribute surface reached by models/mpp.py:115-128
    lin = ours.to_patch_embedding[-1]
    assert isinstance(lin, torch.nn.Linear) and lin.weight.shape == (cfg["dim"], K)
    assert ours.cls_token.shape == (1, 1, cfg["dim"]) and ours.pos_embedding.shape == (1, cfg["num_patches"] + 1, cfg["dim"])
    assert callable(ours.dropout) and callable(ours.transformer)


def test_cpu_input_fails_loudly():
    m = svit.SiT(dim=128, depth=1, heads=2, mlp_dim=128, num_patches=4, num_vertices=5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 4, 4, 5))
    with pytest.raises(ValueError):
        m(torch.randn(2, 4, 4, 6))


@needs_reference
def test_mask_drawing_reproduces_reference_rng_order():
    """draw_masks must consume the torch RNG exactly like models/mpp.py:85-111 (CPU device here)."""
    _, _, mpp_mod = reference_loader.load_reference_models()
    from surface_vision_transformers_b200.mpp import draw_masks
    b, n, k = 6, 20, 12
    mask_prob, replace_prob, swap_prob = 0.5, 0.8, 0.02
    batch = torch.zeros(b, n, k)
    torch.manual_seed(123)
    ref_mask = mpp_mod.get_mask_from_prob(batch, mask_prob)
    rp = mpp_mod.prob_mask_like(batch, swap_prob / (1 - replace_prob))
    ref_swap = ref_mask * (rp == True)  # noqa: E712
    ref_src = torch.randint(0, n, (b, n))
    ref_rep = (ref_mask * mpp_mod.prob_mask_like(batch, replace_prob)) == True  # noqa: E712
    torch.manual_seed(123)
    mask, swap_sel, swap_src, replace_sel = draw_masks(b, n, k, torch.device("cpu"), mask_prob, replace_prob, swap_prob)
    assert torch.equal(mask, ref_mask) and torch.equal(swap_sel, ref_swap)
    assert torch.equal(swap_src, ref_src) and torch.equal(replace_sel, ref_rep)


def test_fused_optimizers_generic_path_matches_torch():
    """Parameters that do not belong to a B200 module take the generic path: must equal torch.optim exactly."""
    torch.manual_seed(0)
    w1 = torch.nn.Parameter(torch.randn(7, 5))
    w2 = torch.nn.Parameter(w1.detach().clone())
    a = svit.FusedAdamW([w1], lr=1e-2, weight_decay=0.1)
    b = torch.optim.AdamW([w2], lr=1e-2, weight_decay=0.1)
    for i in range(4):
        g = torch.randn(7, 5)
        w1.grad, w2.grad = g.clone(), g.clone()
        a.step(); b.step()
    assert torch.allclose(w1, w2, rtol=1e-6, atol=1e-7)
    w3 = torch.nn.Parameter(w1.detach().clone())
    w4 = torch.nn.Parameter(w1.detach().clone())
    c = svit.FusedSGD([w3], lr=1e-2, momentum=0.9, weight_decay=0.01)
    d = torch.optim.SGD([w4], lr=1e-2, momentum=0.9, weight_decay=0.01)
    for i in range(4):
        g = torch.randn(7, 5)
        w3.grad, w4.grad = g.clone(), g.clone()
        c.step(); d.step()
    assert torch.allclose(w3, w4, rtol=1e-6, atol=1e-7)
    assert a.param_groups[0]["lr"] == 1e-2 and "state" in a.state_dict()


def test_stage_segments_partition_flat_buffer():
    m = svit.SiT(dim=128, depth=3, heads=2, mlp_dim=128, num_patches=4, num_vertices=5)
    segs = [m.stage_segment(s) for s in [m.depth] + list(range(m.depth - 1, -1, -1)) + [-1]]
    covered = sorted(segs)
    pos = 0
    for a, n in covered:
        assert a == pos
        pos += n
    assert pos == m._flat.numel()


_DDP_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["SVIT_ROOT"])
import torch, torch.distributed as dist
from surface_vision_transformers_b200.ddp import FlatGradReducer
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["SVIT_PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
G = torch.arange(1000, dtype=torch.float32) * (rank + 1)
r = FlatGradReducer()
for start, n in [(700, 300), (300, 400), (0, 300)]:      # reverse-layer order, like backward
    r.reduce_range(G, start, n)
r.finish()
expect = torch.arange(1000, dtype=torch.float32) * 1.5    # mean of 1x and 2x
assert torch.allclose(G, expect), (G[:5], expect[:5])
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_flat_grad_reducer_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_DDP_WORKER)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), SVIT_PORT=str(port), SVIT_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert "ok" in out
