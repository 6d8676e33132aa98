"""CPU tests of the host side: the C-ABI library loads and exports every symbol of include/svit_b200.h, the
parameter layout / state_dict contract matches the reference, host-side mask drawing reproduces the reference's
RNG order, and the data-parallel gradient reducer works across 2 gloo ranks.  No kernel is launched here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
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
    with pytest.raises(ValueError):
        svit.SiT(dim=128, depth=1, heads=2, mlp_dim=128, dropout=1.0)
    with pytest.raises(ValueError):
        svit.SiT(dim=128, depth=1, heads=2, mlp_dim=128, emb_dropout=-0.1)
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
# DataParallel's three schedules, driven by the engine's callback sequence (head, then per layer: window, layer final; patch
# embedding; "fully enqueued"): every element is averaged exactly once, and the windowed mode issues depth collectives
import surface_vision_transformers_b200 as svit
from surface_vision_transformers_b200.ddp import DataParallel, STAGE_WINDOW
torch.manual_seed(0)
sit = svit.SiT(dim=128, depth=3, heads=2, mlp_dim=128, num_patches=4, num_vertices=5)
for mode, calls in (("window", 3), ("none", 1), ("range", 5)):
    dp = DataParallel(sit, overlap=mode)
    dp.min_window_numel = 1000
    n = sit._flat.numel()
    G = torch.arange(n, dtype=torch.float32) * (rank + 1)
    issued = []
    orig = dp.reducer.reduce_range
    dp.reducer.reduce_range = lambda G_, a, k: (issued.append((a, k)), orig(G_, a, k))
    dp._on_stage(sit, sit.depth, G)
    for l in range(sit.depth - 1, -1, -1):
        dp._on_stage(sit, STAGE_WINDOW + l, G)
        dp._on_stage(sit, l, G)
    dp._on_stage(sit, -1, G)
    dp._on_stage(sit, None, G)
    assert torch.allclose(G, torch.arange(n, dtype=torch.float32) * 1.5), mode
    assert len(issued) == calls and sum(k for _, k in issued) == n, (mode, issued)
    if mode == "window":   # head + last layer first (at the window of layer depth-2), layer 0 + patch embedding last
        assert issued[0][0] + issued[0][1] == n and issued[-1][0] == 0, issued
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


# ------------------------------------------------------------------------------------------- formats
def test_patched_npy_dataset_reads_reference_format(tmp_path):
    """{split}_data.npy is float64 (2S, C, N, V) (tools/preprocessing.py:99); the reader casts to float32 exactly like
    train.py:107 and hands out batches in file order (no shuffle) or as a rank-disjoint permutation (shuffle)."""
    import numpy as np
    rng = np.random.default_rng(0)
    data = rng.standard_normal((10, 4, 20, 15))              # float64
    labels = rng.uniform(26, 45, size=10)
    np.save(tmp_path / "train_data.npy", data)
    np.save(tmp_path / "train_labels.npy", labels)
    ds = svit.PatchedNpyDataset(str(tmp_path), "train", pin=False)
    assert len(ds) == 10 and ds.shape == (10, 4, 20, 15) and ds.data.dtype == torch.float32
    assert torch.equal(ds.data, torch.from_numpy(data).float()) and torch.equal(ds.labels, torch.from_numpy(labels).float())
    got = list(ds.batches(4))
    assert [b[0].shape[0] for b in got] == [4, 4, 2]
    assert torch.equal(torch.cat([b[0] for b in got]), ds.data) and torch.equal(torch.cat([b[1] for b in got]), ds.labels)
    assert [b[0].shape[0] for b in ds.batches(4, drop_last=True)] == [4, 4]
    g = torch.Generator().manual_seed(5)
    r0 = list(ds.batches(3, shuffle=True, generator=g, rank=0, world=2))
    g = torch.Generator().manual_seed(5)
    r1 = list(ds.batches(3, shuffle=True, generator=g, rank=1, world=2))
    seen = torch.cat([b[1] for b in r0] + [b[1] for b in r1])
    assert sorted(seen.tolist()) == sorted(ds.labels.tolist())          # disjoint cover of the epoch
    perm = torch.randperm(10, generator=torch.Generator().manual_seed(5))
    assert torch.equal(torch.cat([b[0] for b in r0]), ds.data[perm[0::2]])


def _timm_like_state_dict(dim, depth, mlp, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {"norm.weight": torch.randn(dim, generator=g), "norm.bias": torch.randn(dim, generator=g),
          "cls_token": torch.randn(1, 1, dim, generator=g), "pos_embed": torch.randn(1, 197, dim, generator=g)}
    for i in range(depth):
        for k, shape in [("norm1.weight", (dim,)), ("norm1.bias", (dim,)), ("norm2.weight", (dim,)), ("norm2.bias", (dim,)),
                         ("attn.qkv.weight", (3 * dim, dim)), ("attn.qkv.bias", (3 * dim,)),
                         ("attn.proj.weight", (dim, dim)), ("attn.proj.bias", (dim,)),
                         ("mlp.fc1.weight", (mlp, dim)), ("mlp.fc1.bias", (mlp,)),
                         ("mlp.fc2.weight", (dim, mlp)), ("mlp.fc2.bias", (dim,))]:
            sd[f"blocks.{i}.{k}"] = torch.randn(*shape, generator=g)
    return sd


def test_load_weights_imagenet_remap():
    """timm ViT -> SiT remap (utils/utils.py:11-35): every encoder tensor and mlp_head.0 replaced, patch embedding /
    cls / pos / mlp_head.1 untouched, result loads strictly into the B200 SiT."""
    cfg = dict(dim=192, depth=3, heads=3, mlp_dim=768, num_patches=20, num_vertices=15)
    model = svit.SiT(**cfg)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    timm = _timm_like_state_dict(192, 3, 768)
    sd = svit.load_weights_imagenet(model.state_dict(), timm, 3)
    model.load_state_dict(sd, strict=True)
    after = model.state_dict()
    assert torch.equal(after["mlp_head.0.weight"], timm["norm.weight"])
    assert torch.equal(after["transformer.layers.2.0.fn.to_qkv.weight"], timm["blocks.2.attn.qkv.weight"])
    assert torch.equal(after["transformer.layers.1.1.fn.net.3.bias"], timm["blocks.1.mlp.fc2.bias"])
    for k in ("pos_embedding", "cls_token", "to_patch_embedding.1.weight", "to_patch_embedding.1.bias", "mlp_head.1.weight"):
        assert torch.equal(after[k], before[k]), k
    changed = [k for k in after if not torch.equal(after[k], before[k])]
    assert len(changed) == 2 + 11 * 3


@needs_reference
def test_load_weights_imagenet_equals_reference_function():
    """Pinned against the reference's own utils/utils.py::load_weights_imagenet (its nibabel import is stubbed: the
    package is absent offline and unrelated to this function)."""
    import importlib.util
    import types
    sys.modules.setdefault("nibabel", types.ModuleType("nibabel"))
    spec = importlib.util.spec_from_file_location("_svit_reference_utils",
                                                  os.path.join(reference_loader.REFERENCE_ROOT, "utils", "utils.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cfg = dict(dim=192, depth=2, heads=3, mlp_dim=768, num_patches=20, num_vertices=15)
    timm = _timm_like_state_dict(192, 2, 768, seed=3)
    base = OracleSiT(**cfg).state_dict()
    a = ref.load_weights_imagenet({k: v.clone() for k, v in base.items()}, timm, 2)
    b = svit.load_weights_imagenet({k: v.clone() for k, v in base.items()}, timm, 2)
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_load_ssl_checkpoint_unwraps_pretrain_format(tmp_path):
    """tools/pretrain.py:378-389 saves {'model_state_dict': ...}; train.py:216 passes the wrapper to load_state_dict
    (a silent no-op).  load_ssl_checkpoint unwraps it, also from the masked_patch_pretraining key space."""
    cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=20, num_vertices=15)
    src = svit.SiT(**cfg)
    dst = svit.SiT(**cfg)
    path = tmp_path / "encoder-best.pt"
    torch.save({"epoch": 3, "model_state_dict": src.state_dict(), "loss": 0.1}, path)
    missing, unexpected = svit.load_ssl_checkpoint(dst, str(path))
    assert not missing and not unexpected
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
    wrapped = {"transformer." + k: v for k, v in src.state_dict().items()}
    wrapped.update({"to_original.weight": torch.zeros(60, 128), "to_original.bias": torch.zeros(60), "mask_token": torch.zeros(1, 1, 60)})
    dst2 = svit.SiT(**cfg)
    missing, unexpected = svit.load_ssl_checkpoint(dst2, {"model_state_dict": wrapped})
    assert not missing and not unexpected
    assert torch.equal(dst2.state_dict()["cls_token"], src.state_dict()["cls_token"])


# ------------------------------------------------------------------------------------------- dropout (8(f)-4)
def test_philox_restatement_matches_random123_known_answers():
    """oracle/dropout.py's Philox4x32-10 against the Random123 known-answer vectors (kat_vectors: zeros, ones, pi)."""
    from oracle.dropout import philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(w) for w in philox4x32_10(*ctr, *key)) == want


def test_dropout_oracle_masks_and_module_semantics():
    from oracle.dropout import MaskedDropout, SITE_EMB, install, keep_mask
    from oracle.sit_oracle import OracleSiT
    n = 1_000_003
    for p in (0.1, 0.5):
        m = keep_mask(n, p, seed=1234, offset=5, site=7)
        assert m.shape == (n,) and abs(m.mean() - (1 - p)) < 4 * (p * (1 - p) / n) ** 0.5
        assert np.array_equal(m, keep_mask(n, p, 1234, 5, 7))                   # pure function of its arguments
        assert not np.array_equal(m, keep_mask(n, p, 1234, 6, 7))               # another step
        assert not np.array_equal(m, keep_mask(n, p, 1234, 5, 8))               # another site
        assert np.array_equal(m[:1001], keep_mask(1001, p, 1234, 5, 7))         # prefix-stable, ragged tail
    assert keep_mask(4097, 0.0, 1, 2, 3).all()
    # independent across elements: lag-1 agreement of a Bernoulli(0.5) stream is 0.5
    m = keep_mask(n, 0.5, 99, 0, SITE_EMB)
    assert abs((m[1:] == m[:-1]).mean() - 0.5) < 3e-3
    d = MaskedDropout(0.25, 1, 0, 3)
    x = torch.randn(7, 33)
    y = d(x)
    kept = y != 0
    assert torch.allclose(y[kept], x[kept] / 0.75) and 0.5 < kept.float().mean() < 0.95
    d.eval()
    assert d(x) is x
    model = install(OracleSiT(dim=64, depth=2, heads=2, mlp_dim=64, num_patches=4, num_vertices=3), 0.1, 0.2, 5, 0)
    sites = [mod.site for mod in model.modules() if isinstance(mod, MaskedDropout)]
    assert sorted(sites) == [0, 1, 2, 4, 5, 6, SITE_EMB]
    assert not any(isinstance(mod, torch.nn.Dropout) for mod in model.modules())


def test_dropout_state_follows_module_mode():
    """Like nn.Dropout: active only in .train(); every training forward advances the offset; eval draws nothing."""
    m = svit.SiT(dim=128, depth=1, heads=2, mlp_dim=128, num_patches=4, num_vertices=3, dropout=0.1, emb_dropout=0.2)
    m.set_dropout_seed(42, step=10)
    assert m._next_dropout_state() == (0.1, 0.2, 42, 10)
    assert m._next_dropout_state(emb=False) == (0.1, 0.0, 42, 11)
    m.eval()
    assert m._next_dropout_state() == (0.0, 0.0, 0, 0)
    m.train()
    assert m._next_dropout_state() == (0.1, 0.2, 42, 12)
    plain = svit.SiT(dim=128, depth=1, heads=2, mlp_dim=128, num_patches=4, num_vertices=3)
    assert plain._next_dropout_state() == (0.0, 0.0, 0, 0) and plain._drop_step == 0
    lib = _lib.load()
    assert lib.svit_set_dropout(m._engine, 1.0, 0.0, 0, 0) != 0 and b"[0, 1)" in lib.svit_last_error()


def test_graph_capture_helpers_reject_what_they_cannot_replay():
    """graphs.py validates before touching the GPU: the DataParallel wrapper, foreign modules and dropout > 0 (whose mask
    offset is host state that a replay would not advance) are refused."""
    from surface_vision_transformers_b200 import graphs
    cfg = dict(dim=128, depth=1, heads=2, mlp_dim=128, num_patches=4, num_vertices=3)
    assert graphs._check_model(svit.SiT(**cfg)) is not None
    with pytest.raises(NotImplementedError):
        graphs._check_model(svit.SiT(**cfg, dropout=0.1))
    with pytest.raises(TypeError):
        graphs._check_model(torch.nn.Linear(4, 4))


@pytest.mark.parametrize("n,world,bs", [(65, 2, 16), (66, 4, 16), (64, 2, 16), (7, 8, 2), (1, 2, 4)])
def test_training_shards_are_equal_on_every_rank(n, world, bs):
    """a training pass all-reduces gradients after EVERY batch, so every rank must see the same number
    of batches with the same sizes whatever len(dataset) % world is (padding by wrap-around, like DistributedSampler);
    evaluation passes keep the exact split."""
    from surface_vision_transformers_b200.data import shard_order
    gens = [torch.Generator().manual_seed(3) for _ in range(world)]
    shards = [shard_order(n, True, gens[r], r, world, True) for r in range(world)]
    sizes = [[len(s[i:i + bs]) for i in range(0, len(s), bs)] for s in shards]
    assert all(sz == sizes[0] for sz in sizes), sizes
    covered = set(torch.cat(shards).tolist())
    assert covered == set(range(n))                                   # nobody is dropped
    assert sum(len(s) for s in shards) == -(-n // world) * world       # padded to a multiple of world
    exact = [shard_order(n, False, None, r, world, False) for r in range(world)]
    assert sorted(torch.cat(exact).tolist()) == list(range(n))        # evaluation: every sample exactly once
