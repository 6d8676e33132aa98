"""-m gpu: SiT / MPP forward + backward through the drop-in modules vs the fp32 oracle on identical weights and
inputs, against the committed golden vectors, and -- at the BASELINE.json sizes -- through size-independent
properties.  Tolerance (north_star): 1e-2 relative L2 for bf16 tensor-core arithmetic."""
import copy

import pytest
import torch

from helpers import CASES, load_golden, rel_l2, seeded_input, seeded_state, subsample
from oracle.sit_oracle import OracleMPP, OracleSiT

pytestmark = pytest.mark.gpu

import surface_vision_transformers_b200 as svit  # noqa: E402

TOL = 1e-2
DEV = torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _fp32_oracle():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def global_grad_rel(named_ours, named_ref):
    ref = dict(named_ref)
    num = sum(((p.grad - ref[n].grad).float() ** 2).sum() for n, p in named_ours if p.grad is not None)
    den = sum((ref[n].grad.float() ** 2).sum() for n, p in named_ours if p.grad is not None)
    return (num / den).sqrt().item()


def check_grads(ours, oracle, tol=TOL):
    """EVERY gradient tensor within the north-star tolerance (1e-2 relative L2) of the fp32 oracle; the worst one is
    printed (pytest -s / the failure message) so the margin stays visible."""
    ref = dict(oracle.named_parameters())
    worst = (0.0, None)
    for n, p in ours.named_parameters():
        assert (p.grad is None) == (ref[n].grad is None), n
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), n
            worst = max(worst, (rel_l2(p.grad, ref[n].grad), n))
    glob = global_grad_rel(list(ours.named_parameters()), oracle.named_parameters())
    print(f"gradients: worst tensor {worst[1]} rel-L2 {worst[0]:.2e}, all tensors together {glob:.2e}")
    assert worst[0] < tol, worst
    assert glob < TOL


@pytest.mark.parametrize("name", ["sit_cls", "sit_mean"])
def test_sit_matches_golden_and_oracle(name):
    case = CASES[name]
    cfg = case["cfg"]
    g = load_golden(name)
    oracle = OracleSiT(**cfg)
    sd = seeded_state(oracle, case["seed"])
    oracle.load_state_dict(sd)
    model = svit.SiT(**cfg)
    model.load_state_dict(sd)
    model.to(DEV)
    x, y = seeded_input(cfg, case["batch"], case["seed"])
    pred = model(x.to(DEV))
    assert pred.shape == (case["batch"], 1) and pred.dtype == torch.float32
    loss = torch.nn.functional.mse_loss(pred.squeeze(), y.to(DEV))
    loss.backward()
    # the prediction is a cancelling scalar (three values of ~0.2 summed from O(1) terms of the tiny golden model): 3e-2
    # like the head output in test_sit_vs_oracle_full_configs; every tensor-valued quantity below keeps the 1e-2 bar
    HEAD_TOL = 3 * TOL
    print(f"{name}: prediction rel-L2 vs golden {rel_l2(pred.detach(), g['pred']):.2e}")
    assert rel_l2(pred.detach(), g["pred"]) < HEAD_TOL
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < TOL
    for k, p in model.named_parameters():
        assert not bool(g["grad_none/" + k]) and p.grad is not None
        assert rel_l2(subsample(p.grad.cpu()), g["grad_sub/" + k]) < 2 * TOL, k   # (a SUBSAMPLE of a small tensor of the tiny golden model: noisier than whole tensors)
    # eval / no_grad path (tools/testing.py:76-88), also with batch 1 (bs_val: 1)
    model.eval()
    with torch.no_grad():
        assert rel_l2(model(x.to(DEV)), g["pred"]) < HEAD_TOL
        assert rel_l2(model(x[:1].to(DEV)), g["pred"][:1]) < HEAD_TOL


def test_mpp_matches_golden():
    case = CASES["mpp"]
    cfg = case["cfg"]
    g = load_golden("mpp")
    K = cfg["num_channels"] * cfg["num_vertices"]
    oracle = OracleMPP(OracleSiT(**cfg), cfg["dim"], K, "cpu", channels=cfg["num_channels"], num_vertices=cfg["num_vertices"], **case["mpp"])
    sd = seeded_state(oracle, case["seed"])
    model = svit.SiT(**cfg)
    ssl = svit.masked_patch_pretraining(transformer=model, dim_in=cfg["dim"], dim_out=K, device=DEV, channels=cfg["num_channels"],
                                        num_vertices=cfg["num_vertices"], **case["mpp"])
    ssl.load_state_dict(sd)
    ssl.to(DEV)
    x, _ = seeded_input(cfg, case["batch"], case["seed"])
    masks = tuple(torch.from_numpy(g[k]).to(DEV) for k in ("mask", "swap_sel", "swap_src", "replace_sel"))
    loss, out = ssl(x.to(DEV), masks=masks)
    assert out.shape == (case["batch"], cfg["num_patches"], K) and loss.dim() == 0
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < TOL
    assert rel_l2(out, g["batch_out"]) < TOL
    for k, p in ssl.named_parameters():
        assert bool(g["grad_none/" + k]) == (p.grad is None), k       # mlp_head.* unreachable -> None, like the reference
        if p.grad is not None:
            assert rel_l2(subsample(p.grad.cpu()), g["grad_sub/" + k]) < 2 * TOL, k
    # eval + no_grad keeps masking active (tools/pretrain.py:345-356)
    ssl.eval()
    with torch.no_grad():
        l2, _ = ssl(x.to(DEV), masks=masks)
    assert abs(l2.item() - float(g["loss"])) / float(g["loss"]) < TOL


CONFIGS = {
    # BASELINE.json configs[0]: SiT-tiny ico-2, batch 16
    "C1_tiny_ico2": (dict(dim=192, depth=12, heads=3, mlp_dim=768, num_patches=320, num_vertices=153), 16),
    # configs[2]: SiT-small ico-1 (large patch-embedding GEMM, T = 81)
    "C3_small_ico1": (dict(dim=384, depth=12, heads=6, mlp_dim=1536, num_patches=80, num_vertices=561), 8),
    # configs[1] at a batch the fp32 oracle finishes quickly
    "C2_small_ico2": (dict(dim=384, depth=12, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153), 6),
    # configs[4] architecture, forward/backward parity at a small batch
    "C5_base_ico2": (dict(dim=768, depth=12, heads=12, mlp_dim=3072, num_patches=320, num_vertices=153), 3),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_sit_vs_oracle_full_configs(name):
    cfg, B = CONFIGS[name]
    torch.manual_seed(0)
    oracle = OracleSiT(**cfg).to(DEV)
    model = svit.SiT(**cfg)
    model.load_state_dict(oracle.state_dict())
    model.to(DEV)
    x = torch.randn(B, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
    y = torch.rand(B, device=DEV) * 19 + 26
    out_o = oracle(x)
    torch.nn.functional.mse_loss(out_o.squeeze(), y).backward()
    out_m = model(x)
    torch.nn.functional.mse_loss(out_m.squeeze(), y).backward()
    # encoder output (B,T,D) is the robust forward comparison (the scalar head nearly cancels)
    with torch.no_grad():
        xe = oracle.to_patch_embedding(x)
        xe = torch.cat((oracle.cls_token.expand(B, -1, -1), xe), 1) + oracle.pos_embedding
        assert rel_l2(model.transformer(xe), oracle.transformer(xe)) < TOL
    # the scalar head output nearly cancels (LayerNorm of the cls row, then a dot product): its relative error amplifies
    # the encoder's, so it keeps a 3e-2 bound -- the measured value is printed; every other tensor is held to 1e-2
    head_err = rel_l2(out_m, out_o)
    print(f"{name}: head output rel-L2 {head_err:.2e}")
    assert head_err < 3 * TOL
    check_grads(model, oracle)


def test_mpp_vs_oracle_small_ico2():
    """configs[3]: SiT-small MPP, 50% mask, replace 0.8, swap 0.02 (config/SiT/pretraining/mpp.yml)."""
    cfg, B = CONFIGS["C2_small_ico2"]
    torch.manual_seed(1)
    K = 4 * cfg["num_vertices"]
    kw = dict(mask_prob=0.5, replace_prob=0.8, swap_prob=0.02, channels=4, num_vertices=cfg["num_vertices"])
    oracle = OracleMPP(OracleSiT(**cfg), cfg["dim"], K, DEV, **kw).to(DEV)
    ssl = svit.masked_patch_pretraining(transformer=svit.SiT(**cfg), dim_in=cfg["dim"], dim_out=K, device=DEV, **kw)
    ssl.load_state_dict(oracle.state_dict())
    ssl.to(DEV)
    x = torch.randn(B, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
    from surface_vision_transformers_b200.mpp import draw_masks
    masks = draw_masks(B, cfg["num_patches"], K, DEV, 0.5, 0.8, 0.02)
    assert (masks[0].sum(1) == 160).all()
    lo, oo = oracle(x, masks=masks)
    lo.backward()
    lm, om = ssl(x, masks=masks)
    lm.backward()
    assert abs(lm.item() - lo.item()) / lo.item() < TOL
    assert rel_l2(om, oo) < TOL
    check_grads(ssl, oracle)
    # seeded RNG path: same device + CPU generator call order as the reference -> same masks -> same loss
    torch.manual_seed(7)
    l1, _ = oracle(x)
    torch.manual_seed(7)
    l2, _ = ssl(x)
    assert abs(l1.item() - l2.item()) / l1.item() < TOL


def test_training_steps_track_oracle():
    """3 optimizer steps: fused AdamW on the flat buffer vs torch.optim.AdamW on the oracle (tools/train.py:280-291)."""
    cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=12, num_vertices=10)
    torch.manual_seed(0)
    oracle = OracleSiT(**cfg).to(DEV)
    model = svit.SiT(**cfg)
    model.load_state_dict(oracle.state_dict())
    model.to(DEV)
    o1 = torch.optim.AdamW(oracle.parameters(), lr=1e-3, weight_decay=0.0)
    o2 = svit.FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.0)
    o3_model = copy.deepcopy(oracle)
    for it in range(3):
        x = torch.randn(5, 4, 12, 10, device=DEV)
        y = torch.rand(5, device=DEV) * 19 + 26
        o1.zero_grad(); o2.zero_grad()
        l1 = torch.nn.functional.mse_loss(oracle(x).squeeze(), y); l1.backward(); o1.step()
        l2 = torch.nn.functional.mse_loss(model(x).squeeze(), y); l2.backward(); o2.step()
        assert abs(l1.item() - l2.item()) / l1.item() < TOL
    moved = sum(((p - q) ** 2).sum() for p, q in zip(oracle.parameters(), o3_model.parameters())).sqrt()
    diff = sum(((p - q) ** 2).sum() for p, q in zip(oracle.parameters(), model.parameters())).sqrt()
    assert moved > 0 and (diff / moved).item() < 0.1      # Adam's sign-like update amplifies bf16 noise: loose bound
    # torch optimizers work on the same parameters too (the YAML default is SGD, tools/train.py:231-236)
    sgd = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9)
    before = model._flat.clone()
    sgd.zero_grad()
    torch.nn.functional.mse_loss(model(x).squeeze(), y).backward()
    sgd.step()
    assert not torch.equal(before, model._flat)
    with torch.no_grad():
        assert torch.isfinite(model(x)).all()


def test_full_size_properties_small_ico2_b64():
    """BASELINE configs[1] at batch 64: size-independent checks -- per-sample independence (batch-split
    consistency of predictions and of the summed gradient), determinism, finite values."""
    cfg = dict(dim=384, depth=12, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153)
    torch.manual_seed(0)
    model = svit.SiT(**cfg).to(DEV)
    x = torch.randn(64, 4, 320, 153, device=DEV)
    y = torch.rand(64, device=DEV) * 19 + 26
    out = model(x)
    ((out.squeeze() - y) ** 2).sum().backward()
    g_full = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    assert torch.isfinite(out).all() and torch.isfinite(g_full).all()
    model.zero_grad()
    outs = []
    for half in (slice(0, 32), slice(32, 64)):
        o = model(x[half])
        ((o.squeeze() - y[half]) ** 2).sum().backward()
        outs.append(o.detach())
    g_sum = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert rel_l2(torch.cat(outs), out.detach()) < 1e-5          # samples are independent: identical arithmetic per row
    assert rel_l2(g_sum, g_full) < 2e-3                          # fp32 atomics / split-K order only
    with torch.no_grad():
        model.eval()
        a = model(x); b = model(x)
        assert torch.equal(a, b)                                  # forward is deterministic
        assert rel_l2(a, out.detach()) < 5e-3                     # eval (fused GELU) vs training path


@pytest.mark.parametrize("cfg,B", [
    (dict(dim=384, depth=3, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153), 5),   # fused Linear + LayerNorm tile
    (dict(dim=192, depth=2, heads=3, mlp_dim=768, num_patches=80, num_vertices=45), 7),      # stand-alone LayerNorm kernels
    (dict(dim=128, depth=1, heads=2, mlp_dim=256, num_patches=20, num_vertices=15), 3),      # the last block is the first
])
def test_cls_pooling_last_block_on_token_zero_only(cfg, B, monkeypatch):
    """Under cls pooling the engine runs the last block's attention for the cls query alone and its row-wise rest on B
    rows (engine.cu cls_last, attention_cls.cu).  Same prediction and same gradient for EVERY parameter as the engine
    that computes all T rows (SVIT_FULL_LAST_LAYER=1, read at svit_create) -- training and inference -- and as the oracle."""
    torch.manual_seed(5)
    oracle = OracleSiT(**cfg).to(DEV)
    x = torch.randn(B, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
    y = torch.rand(B, device=DEV) * 19 + 26
    res = {}
    for name, full in (("cls", "0"), ("full", "1")):
        monkeypatch.setenv("SVIT_FULL_LAST_LAYER", full)
        m = svit.SiT(**cfg)
        m.load_state_dict(oracle.state_dict())
        m.to(DEV)
        out = m(x)
        torch.nn.functional.mse_loss(out.squeeze(), y).backward()
        with torch.no_grad():
            ev = m.eval()(x).clone()
        res[name] = (out.detach().clone(), {n: p.grad.clone() for n, p in m.named_parameters()}, ev)
    lo = torch.nn.functional.mse_loss(oracle(x).squeeze(), y)
    lo.backward()
    go = {n: p.grad for n, p in oracle.named_parameters()}
    assert rel_l2(res["cls"][0], res["full"][0]) < 3e-3 and rel_l2(res["cls"][2], res["full"][2]) < 3e-3
    worst = max((rel_l2(res["cls"][1][n], res["full"][1][n]), n) for n in go)
    worst_o = max((rel_l2(res["cls"][1][n], go[n]), n) for n in go)
    print(f"cls-only last block vs all rows: worst tensor {worst[1]} {worst[0]:.2e}; vs oracle: {worst_o[1]} {worst_o[0]:.2e}")
    assert worst[0] < TOL and worst_o[0] < TOL
    for n in go:                                   # no gradient went missing: same support
        assert torch.isfinite(res["cls"][1][n]).all() and (res["cls"][1][n].abs().sum() > 0) == (go[n].abs().sum() > 0), n


@pytest.mark.parametrize("cfg,B,pool", [
    (dict(dim=384, depth=3, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153), 6, "cls"),   # B-row last block + two full layers
    (dict(dim=384, depth=2, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153), 4, "mean"),  # every layer full size
    (dict(dim=192, depth=3, heads=3, mlp_dim=768, num_patches=80, num_vertices=45), 9, "cls"),      # stand-alone LayerNorm kernels
])
def test_side_stream_weight_gradients_equal_single_stream(cfg, B, pool, monkeypatch):
    """engine.cu runs the weight-gradient GEMMs on a side stream next to the LayerNorm backward kernels (DESIGN 3d;
    SVIT_WGRAD_OVERLAP, read at svit_create: 0 = one stream, 1 = the adjacent wgrad only, 2 = default).  The kernels and their
    operands are the same in every mode, so every gradient must agree with the one-stream engine up to the order of the fp32
    reductions (red.global.add / atomicAdd) -- a missing event dependency would show up as a stale or half-written operand.
    Repeated backward passes (the scratch buffers are reused from layer to layer and from pass to pass) must agree as well."""
    torch.manual_seed(11)
    oracle = OracleSiT(pool=pool, **cfg).to(DEV)
    x = torch.randn(B, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
    y = torch.rand(B, device=DEV) * 19 + 26
    grads = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("SVIT_WGRAD_OVERLAP", mode)
        m = svit.SiT(pool=pool, **cfg)
        m.load_state_dict(oracle.state_dict())
        m.to(DEV)
        runs = []
        for _ in range(3):
            m.zero_grad(set_to_none=True)
            torch.nn.functional.mse_loss(m(x).squeeze(), y).backward()
            runs.append({n: p.grad.clone() for n, p in m.named_parameters()})
        torch.cuda.synchronize()
        for r in runs[1:]:
            for n in r:
                assert rel_l2(r[n], runs[0][n]) < 1e-5, (mode, n)
        grads[mode] = runs[0]
    for mode in ("1", "2"):
        worst = max((rel_l2(grads[mode][n], grads["0"][n]), n) for n in grads["0"])
        print(f"SVIT_WGRAD_OVERLAP={mode} vs 0: worst tensor {worst[1]} {worst[0]:.2e}")
        assert worst[0] < 1e-5, (mode, worst)


def test_side_stream_weight_gradients_mpp(monkeypatch):
    """The same through the MPP module (svit_mpp_backward: every layer full size, decoder in front of the encoder backward)."""
    cfg = dict(dim=384, depth=2, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153)
    torch.manual_seed(12)
    ref = svit.SiT(**cfg)
    x = torch.randn(3, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
    K = 4 * cfg["num_vertices"]
    grads = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("SVIT_WGRAD_OVERLAP", mode)
        sit = svit.SiT(**cfg)
        sit.load_state_dict(ref.state_dict())
        torch.manual_seed(13)
        ssl = svit.masked_patch_pretraining(transformer=sit, dim_in=cfg["dim"], dim_out=K, device=DEV, mask_prob=0.5,
                                            replace_prob=0.8, swap_prob=0.02, channels=4,
                                            num_vertices=cfg["num_vertices"]).to(DEV)
        torch.manual_seed(14)                      # same masks in both modes (device and CPU generators, Appendix B order)
        loss, _ = ssl(x)
        loss.backward()
        torch.cuda.synchronize()
        grads[mode] = {n: p.grad.clone() for n, p in ssl.named_parameters() if p.grad is not None}
    assert grads["0"].keys() == grads["2"].keys()
    worst = max((rel_l2(grads["2"][n], grads["0"][n]), n) for n in grads["0"])
    print(f"MPP, SVIT_WGRAD_OVERLAP=2 vs 0: worst tensor {worst[1]} {worst[0]:.2e}")
    assert worst[0] < 1e-5, worst


def test_raw_mesh_ingestion_matches_prepatched():
    """raw ico-6 mesh + on-device gather/z-score == pre-patched input through the same network."""
    cfg = dict(dim=192, depth=2, heads=3, mlp_dim=768, num_patches=320, num_vertices=153)
    torch.manual_seed(0)
    model = svit.SiT(**cfg).to(DEV).eval()
    table = svit.load_index_table(2, DEV)
    mesh = torch.randn(3, 4, 40962, device=DEV) * 2 + 1
    mean = torch.tensor([1.15, 0.037, 1.0, 0.07], device=DEV)
    std = torch.tensor([0.41, 0.19, 0.39, 4.05], device=DEV)
    patched = svit.gather_patches((mesh - mean.view(1, 4, 1)) / std.view(1, 4, 1), table)
    with torch.no_grad():
        a = model(patched)
        b = model.forward_mesh(mesh, table, mean, std)
    assert rel_l2(b, a) < 2e-3


@pytest.mark.parametrize("depth,nbatches", [(2, 7), (3, 5), (2, 1)])
def test_device_prefetcher_order_and_content(depth, nbatches):
    """Batches come out bit-identical and in order while a slow consumer keeps the compute stream busy (the copy of
    batch i+1 runs on the side stream meanwhile and must never overwrite a batch that is still being read)."""
    torch.manual_seed(0)
    host = [(torch.randn(64, 4, 20, 15).pin_memory(), torch.rand(64).pin_memory()) for _ in range(nbatches)]
    w = torch.randn(2048, 2048, device=DEV)
    seen = []
    for x, y in svit.DevicePrefetcher(iter(host), DEV, depth=depth):
        for _ in range(10):           # slow consumer
            w = torch.tanh(w @ w * 1e-3)
        seen.append((x.clone(), y.clone()))
    torch.cuda.synchronize()
    assert len(seen) == nbatches
    for (x, y), (hx, hy) in zip(seen, host):
        assert torch.equal(x.cpu(), hx) and torch.equal(y.cpu(), hy)


CHECK_TOL = 1e-4   # north star: "... tightening to 1e-4 in an fp32-accumulate check mode"


@pytest.mark.parametrize("name,pool", [("C1_tiny_ico2", "cls"), ("C3_small_ico1", "mean"), ("C2_small_ico2", "cls")])
def test_fp32_check_mode_matches_oracle_to_1e4(name, pool):
    """set_check_mode(True): the same engine orchestration with every operand in fp32 on the CUDA cores.  Forward
    outputs, the encoder output and EVERY parameter gradient (per tensor, not just globally) within 1e-4 relative L2 of
    the fp32 oracle on identical weights and inputs; switching back restores the bf16 tensor-core path."""
    cfg, B = CONFIGS[name]
    cfg = dict(cfg, depth=4, pool=pool)       # 4 blocks keep the un-tuned fp32 kernels quick
    B = min(B, 4)
    torch.manual_seed(1)
    oracle = OracleSiT(**cfg).to(DEV)
    model = svit.SiT(**cfg)
    model.load_state_dict(oracle.state_dict())
    model.to(DEV).set_check_mode(True)
    assert model.check_mode
    x = torch.randn(B, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
    y = torch.rand(B, device=DEV) * 19 + 26
    out_o = oracle(x)
    torch.nn.functional.mse_loss(out_o.squeeze(), y).backward()
    out_m = model(x)
    torch.nn.functional.mse_loss(out_m.squeeze(), y).backward()
    assert rel_l2(out_m, out_o) < CHECK_TOL
    ref = dict(oracle.named_parameters())
    worst = max((rel_l2(p.grad, ref[n].grad), n) for n, p in model.named_parameters())
    assert worst[0] < CHECK_TOL, worst
    with torch.no_grad():
        xe = oracle.to_patch_embedding(x)
        xe = torch.cat((oracle.cls_token.expand(B, -1, -1), xe), 1) + oracle.pos_embedding
        assert rel_l2(model.transformer(xe), oracle.transformer(xe)) < CHECK_TOL
        # back to the tensor-core path: same call, bf16 tolerance, visibly different bits
        model.set_check_mode(False)
        out_bf16 = model(x)
        assert not model.check_mode and rel_l2(out_bf16, out_o) < 3 * TOL


def test_fp32_check_mode_mpp_matches_oracle_to_1e4():
    """Check mode through the MPP module: corruption, decoder, masked L2 loss and every gradient (encoder, patch
    embedding, to_original, mask_token) within 1e-4 of the fp32 oracle on the same masks."""
    cfg = dict(dim=192, depth=3, heads=3, mlp_dim=768, num_patches=80, num_vertices=45)
    B, K = 3, 4 * 45
    torch.manual_seed(3)
    kw = dict(mask_prob=0.5, replace_prob=0.8, swap_prob=0.1, channels=4, num_vertices=cfg["num_vertices"])
    oracle = OracleMPP(OracleSiT(**cfg), cfg["dim"], K, DEV, **kw).to(DEV)
    ssl = svit.masked_patch_pretraining(transformer=svit.SiT(**cfg), dim_in=cfg["dim"], dim_out=K, device=DEV, **kw)
    ssl.load_state_dict(oracle.state_dict())
    ssl.to(DEV)
    ssl.transformer.set_check_mode(True)
    x = torch.randn(B, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
    from surface_vision_transformers_b200.mpp import draw_masks
    masks = draw_masks(B, cfg["num_patches"], K, DEV, 0.5, 0.8, 0.1)
    lo, oo = oracle(x, masks=masks)
    lo.backward()
    lm, om = ssl(x, masks=masks)
    lm.backward()
    assert abs(lm.item() - lo.item()) / lo.item() < CHECK_TOL
    assert rel_l2(om, oo) < CHECK_TOL
    ref = dict(oracle.named_parameters())
    for n, p in ssl.named_parameters():
        if ref[n].grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
        else:
            assert rel_l2(p.grad, ref[n].grad) < CHECK_TOL, n


# ------------------------------------------------------------------------------------------- SURVEY 8(f) rows
@pytest.mark.parametrize("sub_ico", [1, 2])
def test_preprocess_meshes_bit_exact_vs_reference_loop(sub_ico):
    """8(f)-1: on-device z-score + gather + L/R re-ordering == the numpy restatement of tools/preprocessing.py:72-84
    (float64 z-score like the reference, cast to float32 like tools/train.py:107), bit for bit."""
    import numpy as np
    from oracle import gather_oracle
    rs = np.random.RandomState(7 + sub_ico)
    hemis = (rs.standard_normal((6, 4, 40962)) * 2 + 1).astype(np.float32).astype(np.float64)   # L0,R0,L1,R1,L2,R2
    means = np.array([1.15, 0.037, 1.0, 0.07]).reshape(1, 4, 1)
    stds = np.array([0.41, 0.19, 0.39, 4.05]).reshape(1, 4, 1)
    table = svit.load_index_table(sub_ico)
    ref = gather_oracle.preprocessing_layout(gather_oracle.zscore(hemis, means, stds), table.numpy()).astype(np.float32)
    got = svit.preprocess_meshes(torch.from_numpy(hemis).to(DEV), means.reshape(-1), stds.reshape(-1), table.to(DEV))
    assert got.dtype == torch.float32 and tuple(got.shape) == ref.shape
    assert torch.equal(got.cpu(), torch.from_numpy(ref))


def test_fit_tracks_reference_style_loop(tmp_path):
    """8(f)-2: fit() (prefetched batches, device-side epoch statistics, best-MAE checkpoint) against a plain
    tools/train.py-style loop on the fp32 oracle with the same data order: same per-epoch loss / MAE within the bf16
    tolerance, validation bookkeeping and the checkpoint file as in train.py:339-363."""
    import numpy as np
    cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=12, num_vertices=10)
    rng = np.random.default_rng(0)
    for split, n in (("train", 24), ("validation", 10)):
        np.save(tmp_path / f"{split}_data.npy", rng.standard_normal((n, 4, 12, 10)))
        np.save(tmp_path / f"{split}_labels.npy", rng.uniform(26, 45, size=n))
    train = svit.PatchedNpyDataset(str(tmp_path), "train")
    val = svit.PatchedNpyDataset(str(tmp_path), "validation")
    torch.manual_seed(0)
    oracle = OracleSiT(**cfg).to(DEV)
    model = svit.SiT(**cfg)
    model.load_state_dict(oracle.state_dict())
    model.to(DEV)
    lines = []
    res = svit.fit(model, svit.FusedAdamW(model.parameters(), lr=3e-4, weight_decay=0.0), train, val, epochs=3, batch_size=8,
                   val_epoch=1, device=DEV, save_dir=str(tmp_path / "out"), save_ckpt=True, seed=11, log=lines.append)
    # reference-style loop on the oracle with the same permutations
    opt = torch.optim.AdamW(oracle.parameters(), lr=3e-4, weight_decay=0.0)
    gen = torch.Generator()
    ref_loss, ref_mae, ref_vmae = [], [], []
    for epoch in range(3):
        gen.manual_seed(11 + epoch)
        oracle.train()
        run, preds, targs = 0.0, [], []
        batches = list(train.batches(8, shuffle=True, generator=gen))
        for x, y in batches:
            x, y = x.to(DEV), y.to(DEV)
            opt.zero_grad()
            out = oracle(x)
            loss = torch.nn.functional.mse_loss(out.squeeze(), y)
            loss.backward()
            opt.step()
            run += loss.item()
            preds.append(out.reshape(-1).detach().cpu()); targs.append(y.cpu())
        ref_loss.append(run / len(batches))
        ref_mae.append((torch.cat(targs) - torch.cat(preds)).abs().mean().item())
        oracle.eval()
        with torch.no_grad():
            vp_ = torch.cat([oracle(x.to(DEV)).reshape(-1).cpu() for x, _ in val.batches(8)])
        ref_vmae.append((val.labels - vp_).abs().mean().item())
    h = res["history"]
    for e in range(3):
        assert abs(h["train_loss"][e] - ref_loss[e]) / ref_loss[e] < 3 * TOL, (e, h["train_loss"][e], ref_loss[e])
        assert abs(h["train_mae"][e] - ref_mae[e]) / ref_mae[e] < 3 * TOL
        assert h["val_mae"][e][0] == e + 1 and abs(h["val_mae"][e][1] - ref_vmae[e]) / ref_vmae[e] < 3 * TOL
    assert res["best_epoch"] == int(np.argmin(ref_vmae)) + 1
    assert abs(res["best_mae"] - min(ref_vmae)) / min(ref_vmae) < 3 * TOL
    assert len(lines) == 6 and lines[1].startswith("| Validation | Epoch - 1 |")
    ckpt = torch.load(tmp_path / "out" / "checkpoint.pth")
    assert set(ckpt) == set(oracle.state_dict())
    saved = torch.load(tmp_path / "out" / "preds_test.pt")
    assert saved["preds"].shape == (10,) and torch.equal(saved["targets"], val.labels)


def test_fit_mpp_tracks_reference_style_loop(tmp_path):
    """8(f)-2, pre-training half: fit_mpp() against a plain tools/pretrain.py-style loop (pretrain.py:303-389) on the
    fp32 oracle.  Both consume torch's RNG in the reference's order, so they draw the same masks; per-epoch train and
    validation losses agree within the bf16 tolerance, the best epoch is the same, and the two checkpoint files have
    the reference's format -- load_ssl_checkpoint() reads the encoder back for fine-tuning."""
    import numpy as np
    cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=20, num_vertices=15)
    K = 4 * 15
    rng = np.random.default_rng(1)
    for split, n in (("train", 24), ("validation", 8)):
        np.save(tmp_path / f"{split}_data.npy", rng.standard_normal((n, 4, 20, 15)))
        np.save(tmp_path / f"{split}_labels.npy", rng.uniform(26, 45, size=n))
    train = svit.PatchedNpyDataset(str(tmp_path), "train")
    val = svit.PatchedNpyDataset(str(tmp_path), "validation")
    kw = dict(mask_prob=0.5, replace_prob=0.8, swap_prob=0.02, channels=4, num_vertices=15)
    torch.manual_seed(2)
    oracle = OracleMPP(OracleSiT(**cfg), cfg["dim"], K, DEV, **kw).to(DEV)
    ssl = svit.masked_patch_pretraining(transformer=svit.SiT(**cfg), dim_in=cfg["dim"], dim_out=K, device=DEV, **kw)
    ssl.load_state_dict(oracle.state_dict())
    ssl.to(DEV)
    epochs, bs = 3, 8
    # reference-style loop on the oracle
    opt = torch.optim.AdamW(oracle.parameters(), lr=3e-4, weight_decay=0.0)
    gen = torch.Generator()
    torch.manual_seed(99)
    ref_train, ref_val = [], []
    for epoch in range(epochs):
        gen.manual_seed(5 + epoch)
        oracle.train()
        run, nb = 0.0, 0
        for x, _ in train.batches(bs, shuffle=True, generator=gen):
            opt.zero_grad()
            loss, _ = oracle(x.to(DEV))
            loss.backward()
            opt.step()
            run += loss.item(); nb += 1
        ref_train.append(run / nb)
        oracle.eval()
        with torch.no_grad():
            vl = [oracle(x.to(DEV))[0].item() for x, _ in val.batches(bs)]
        ref_val.append(sum(vl) / len(vl))
    torch.manual_seed(99)
    opt2 = svit.FusedAdamW(ssl.parameters(), lr=3e-4, weight_decay=0.0)
    res = svit.fit_mpp(ssl, opt2, train, val, epochs=epochs, batch_size=bs, val_epoch=1, device=DEV,
                       save_dir=str(tmp_path / "out"), seed=5)
    h = res["history"]
    for a, b in zip(h["train_loss"], ref_train):
        assert abs(a - b) / b < 3 * TOL, (h["train_loss"], ref_train)
    for (e, a), b in zip(h["val_loss"], ref_val):
        assert abs(a - b) / b < 3 * TOL, (h["val_loss"], ref_val)
    assert h["train_loss"][-1] < h["train_loss"][0]
    assert res["best_epoch"] == 1 + min(range(epochs), key=lambda i: h["val_loss"][i][1])   # bookkeeping of pretrain.py:362-365
    assert abs(res["best_val_loss"] - min(ref_val)) / min(ref_val) < 3 * TOL
    ck = torch.load(tmp_path / "out" / "encoder-decoder-best.pt")
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "loss"} and ck["epoch"] == res["best_epoch"]
    assert set(ck["model_state_dict"]) == set(oracle.state_dict())
    enc = torch.load(tmp_path / "out" / "encoder-best.pt")
    assert set(enc["model_state_dict"]) == set(oracle.transformer.state_dict())
    fresh = svit.SiT(**cfg)
    svit.load_ssl_checkpoint(fresh, str(tmp_path / "out" / "encoder-best.pt"))
    for k, v in enc["model_state_dict"].items():
        assert torch.equal(fresh.state_dict()[k].cpu(), v.cpu()), k


def test_empty_batch_returns_empty_prediction():
    """B = 0 (an empty last shard of a sampler): like the reference, an empty (0, num_classes) tensor that still
    back-propagates (zero gradients) instead of an engine error."""
    model = svit.SiT(dim=128, depth=1, heads=2, mlp_dim=128, num_patches=20, num_vertices=15).to(DEV)
    out = model(torch.empty(0, 4, 20, 15, device=DEV))
    assert out.shape == (0, 1) and out.dtype == torch.float32
    out.sum().backward()
    with torch.no_grad():
        assert model.eval()(torch.empty(0, 4, 20, 15, device=DEV)).shape == (0, 1)


def _param_rel(a, b):
    num = sum(((p.detach() - q.detach()).float() ** 2).sum() for p, q in zip(a.parameters(), b.parameters()))
    den = sum((q.detach().float() ** 2).sum() for q in b.parameters())
    return (num / den).sqrt().item()


def test_adamw_step_bookkeeping_is_stream_ordered_and_restorable():
    """The AdamW step counters / bias corrections live on the device: (a) a host that enqueues many steps without ever
    synchronising gets the same parameters as one that synchronises after every step (staging them through a pinned
    buffer did not guarantee that: the host could rewrite it before the copy ran); (b) optimizer.state_dict() ->
    load_state_dict() into a fresh optimizer continues the run (moments, step counts, device table).

    The strict comparison runs in fp32 check mode.  On the bf16 path two runs of the SAME program end either 5e-9 or 3e-5
    apart (scripts/probe_drift.py, probe_drift2.py; gpurun_out/r2l_drift*.log): the fp32 atomics of the weight-gradient
    kernels make step-1 gradients differ by 5e-8, AdamW turns that into last-bit differences of a few parameters, and one
    of them moves an activation of step 2 across a bf16 rounding boundary (forward output 7e-5, gradients 5e-5 apart).
    Replaying step 2 from bit-identical parameters is deterministic (12 of 12), so this is rounding chaos, not a missed
    hand-off; a wrong bias correction on one of the 8 steps would show as >= 5e-3, which the loose bound still catches."""
    cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=20, num_vertices=15)
    torch.manual_seed(21)
    base = svit.SiT(**cfg).to(DEV)
    xs = [torch.randn(8, 4, 20, 15, device=DEV) for _ in range(8)]
    ys = [torch.rand(8, device=DEV) * 19 + 26 for _ in range(8)]

    def fresh(check, state=None):
        m = svit.SiT(**cfg)
        m.load_state_dict(base.state_dict() if state is None else state)
        m.to(DEV)
        m.set_check_mode(check)
        return m, svit.FusedAdamW(m.parameters(), lr=1e-3, weight_decay=0.0)

    def run(model, opt, lo, hi, sync):
        for k in range(lo, hi):
            opt.zero_grad(set_to_none=True)
            torch.nn.functional.mse_loss(model(xs[k]).squeeze(), ys[k]).backward()
            opt.step()
            if sync:
                torch.cuda.synchronize()

    for check, bound in ((True, 1e-6), (False, 1e-3)):
        models = []
        for sync in (False, True):
            m, o = fresh(check)
            run(m, o, 0, 8, sync)
            models.append((m, o))
        torch.cuda.synchronize()
        drift = _param_rel(models[0][0], models[1][0])
        print(f"un-synchronised vs synchronised host after 8 AdamW steps ({'fp32 check mode' if check else 'bf16 path'}): "
              f"parameter rel-L2 {drift:.2e}")
        assert drift < bound, drift
        assert float(models[0][1].state[models[0][0].cls_token]["step"]) == 8.0
        # (b) resume from a checkpoint taken after 4 steps
        m1, o1 = fresh(check)
        run(m1, o1, 0, 4, True)
        ck_m = {k: v.clone() for k, v in m1.state_dict().items()}
        ck_o = copy.deepcopy(o1.state_dict())
        m2, o2 = fresh(check, ck_m)
        o2.load_state_dict(ck_o)
        run(m2, o2, 4, 8, True)
        torch.cuda.synchronize()
        assert _param_rel(m2, models[1][0]) < bound
        assert float(o2.state[m2.cls_token]["step"]) == 8.0


def test_cuda_graph_capture_of_inference_and_training_step():
    """8(f)-2: the whole step replays from a CUDA graph.  Inference: bit-identical to the eager call.  Training (zero_grad
    + forward + MSE + backward + FusedAdamW.step in ONE graph): parameters after warm-up + 5 replays on changing batches
    equal those of the same eager steps (the bias corrections advance on the device inside the graph)."""
    cfg = dict(dim=192, depth=3, heads=3, mlp_dim=768, num_patches=80, num_vertices=45)
    torch.manual_seed(22)
    base = svit.SiT(**cfg).to(DEV)
    B = 8
    xs = [torch.randn(B, 4, 80, 45, device=DEV) for _ in range(6)]
    ys = [torch.rand(B, device=DEV) * 19 + 26 for _ in range(6)]
    base.eval()
    with torch.no_grad():
        want = [base(x).clone() for x in xs[:3]]
    gi = svit.GraphedInference(base, xs[0])
    for x, w in zip(xs[:3], want):
        assert torch.equal(gi(x), w)
    with torch.no_grad():                      # the graph re-derives the weight shadows: it follows weight updates
        base.mlp_head[1].bias.add_(1.0)
        base.transformer.layers[0][1].fn.net[0].weight.mul_(1.5)
        base.mark_weights_dirty()
        assert torch.equal(gi(xs[1]), base(xs[1])) and not torch.equal(gi(xs[1]), want[1])
    crit = lambda out, t: torch.nn.functional.mse_loss(out.squeeze(-1), t)   # noqa: E731
    eager = svit.SiT(**cfg); eager.load_state_dict(base.state_dict()); eager.to(DEV).train()
    graphed = svit.SiT(**cfg); graphed.load_state_dict(base.state_dict()); graphed.to(DEV).train()
    oe = svit.FusedAdamW(eager.parameters(), lr=1e-3, weight_decay=0.0)
    og = svit.FusedAdamW(graphed.parameters(), lr=1e-3, weight_decay=0.0)
    step = svit.GraphedTrainStep(graphed, og, crit, xs[0], ys[0], warmup=3)    # 3 real warm-up steps on batch 0, then the capture
    for _ in range(3):
        oe.zero_grad(set_to_none=True); crit(eager(xs[0]), ys[0]).backward(); oe.step()
    losses = []
    for k in range(1, 6):
        oe.zero_grad(set_to_none=True)
        le = crit(eager(xs[k]), ys[k]); le.backward(); oe.step()
        lg = step(xs[k], ys[k])
        losses.append((le.item(), lg.item()))
    torch.cuda.synchronize()
    for le, lg in losses:
        assert abs(le - lg) / abs(le) < 1e-3, losses
    assert _param_rel(graphed, eager) < 1e-4
    assert float(og.state[graphed.cls_token]["step"]) == float(oe.state[eager.cls_token]["step"]) == 8.0
    og.param_groups[0]["lr"] = 5e-4
    with pytest.raises(RuntimeError):
        step(xs[0], ys[0])
    step.recapture()
    assert torch.isfinite(step(xs[0], ys[0])).all()


# ------------------------------------------------------------------------------------------- dropout > 0 (8(f)-4)
def _dropout_pair(cfg, p, emb_p, seed, step, check_mode):
    """B200 SiT with dropout and the oracle with MaskedDropout modules applying the very same keep decisions."""
    from oracle.dropout import install
    oracle = OracleSiT(**cfg).to(DEV)
    model = svit.SiT(**cfg, dropout=p, emb_dropout=emb_p)
    model.load_state_dict(oracle.state_dict())
    model.to(DEV).set_check_mode(check_mode).set_dropout_seed(seed, step)
    install(oracle, p, emb_p, seed, step)
    return model, oracle


@pytest.mark.parametrize("n,p,seed,offset,site", [(4096, 0.1, 1, 0, 0), (1001, 0.5, 0xDEADBEEFCAFEF00D, 2 ** 40 + 3, 46),
                                                  (3, 0.25, 7, 1, 0xFFFF0000), (82176 * 96 + 1, 0.1, 123, 456, 5),
                                                  (257, 0.0, 9, 9, 9)])
def test_dropout_mask_generator_bit_exact(n, p, seed, offset, site):
    """svit_dropout_mask (the Philox4x32-10 keep decisions every dropout kernel uses) == the numpy restatement."""
    from oracle.dropout import keep_mask
    from surface_vision_transformers_b200 import _lib
    keep = torch.zeros(n, dtype=torch.uint8, device=DEV)
    _lib.check(_lib.load().svit_dropout_mask(_lib.ptr(keep), n, p, seed, offset, site,
                                            _lib.vp(torch.cuda.current_stream().cuda_stream)), "svit_dropout_mask")
    want = keep_mask(n, p, seed, offset, site)
    assert torch.equal(keep.cpu().bool(), torch.from_numpy(want))


@pytest.mark.parametrize("pool", ["cls", "mean"])
def test_dropout_fp32_check_mode_matches_oracle_with_same_masks(pool):
    """dropout 0.1 / emb_dropout 0.2 (the reference's four nn.Dropout sites) in fp32 check mode: output, encoder output
    and every parameter gradient within 1e-4 of the oracle applying the same masks."""
    cfg = dict(dim=192, depth=3, heads=3, mlp_dim=768, num_patches=80, num_vertices=45, pool=pool)
    B = 5
    torch.manual_seed(11)
    model, oracle = _dropout_pair(cfg, 0.1, 0.2, seed=2024, step=3, check_mode=True)
    x = torch.randn(B, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
    y = torch.rand(B, device=DEV) * 19 + 26
    out_o = oracle(x)
    torch.nn.functional.mse_loss(out_o.squeeze(), y).backward()
    out_m = model(x)
    torch.nn.functional.mse_loss(out_m.squeeze(), y).backward()
    assert rel_l2(out_m, out_o) < CHECK_TOL
    ref = dict(oracle.named_parameters())
    worst = max((rel_l2(p.grad, ref[n].grad), n) for n, p in model.named_parameters())
    assert worst[0] < CHECK_TOL, worst
    # the masks matter: the no-dropout forward is far away
    model.eval()
    with torch.no_grad():
        assert rel_l2(model(x), out_o) > 10 * CHECK_TOL


@pytest.mark.parametrize("name", ["C1_tiny_ico2", "C2_small_ico2"])
def test_dropout_bf16_matches_oracle_with_same_masks(name):
    """The tensor-core path with dropout > 0 against the fp32 oracle on the same masks, bf16 tolerance; the encoder
    alone (model.transformer(x), no emb_dropout) too."""
    cfg, B = CONFIGS[name]
    cfg = dict(cfg, depth=6)
    torch.manual_seed(12)
    model, oracle = _dropout_pair(cfg, 0.1, 0.1, seed=77, step=0, check_mode=False)
    x = torch.randn(B, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
    y = torch.rand(B, device=DEV) * 19 + 26
    out_o = oracle(x)
    torch.nn.functional.mse_loss(out_o.squeeze(), y).backward()
    out_m = model(x)                                       # offset 0
    torch.nn.functional.mse_loss(out_m.squeeze(), y).backward()
    # the scalar head output nearly cancels (LayerNorm of the cls row, then a dot product): its relative error amplifies
    # the encoder's, so it keeps a 3e-2 bound -- the measured value is printed; every other tensor is held to 1e-2
    head_err = rel_l2(out_m, out_o)
    print(f"{name}: head output rel-L2 {head_err:.2e}")
    assert head_err < 3 * TOL
    check_grads(model, oracle)
    from oracle.dropout import install
    install(oracle, 0.1, 0.1, 77, 1)                       # the next forward of `model` uses offset 1
    xe = torch.randn(B, cfg["num_patches"] + 1, cfg["dim"], device=DEV, requires_grad=True)
    xo = xe.detach().clone().requires_grad_(True)
    model.zero_grad(), oracle.zero_grad()
    ye = model.transformer(xe)
    yo = oracle.transformer(xo)
    assert rel_l2(ye, yo) < TOL
    w = torch.randn_like(yo)
    (ye * w).sum().backward()
    (yo * w).sum().backward()
    assert rel_l2(xe.grad, xo.grad) < TOL
    enc = [(n, p) for n, p in model.named_parameters() if n.startswith("transformer.")]
    assert global_grad_rel(enc, oracle.named_parameters()) < TOL


def test_dropout_train_eval_semantics_and_statistics():
    """eval() == the dropout-free model bit for bit; train() differs, changes from step to step, is reproducible from
    (seed, step), and drops under no_grad as nn.Dropout does; the mean over many masks approaches the eval output of
    the first dropout site (unbiased 1/(1-p) scaling)."""
    cfg = dict(dim=192, depth=2, heads=3, mlp_dim=768, num_patches=80, num_vertices=45)
    torch.manual_seed(13)
    plain = svit.SiT(**cfg).to(DEV)
    model = svit.SiT(**cfg, dropout=0.2, emb_dropout=0.1)
    model.load_state_dict(plain.state_dict())
    model.to(DEV)
    x = torch.randn(4, 4, 80, 45, device=DEV)
    with torch.no_grad():
        model.eval()
        assert torch.equal(model(x), plain(x))
        model.train().set_dropout_seed(5, 0)
        a, b = model(x), model(x)
        assert not torch.equal(a, b) and not torch.equal(a, plain(x))
        model.set_dropout_seed(5, 0)
        assert torch.equal(model(x), a) and torch.equal(model(x), b)
    # unbiasedness at one site: encoder of depth 1 with only emb_dropout = linear in the mask up to the first LayerNorm
    one = svit.SiT(**dict(cfg, depth=1), emb_dropout=0.3).to(DEV)
    from surface_vision_transformers_b200 import _lib
    n = 1 << 20
    keep = torch.zeros(n, dtype=torch.uint8, device=DEV)
    acc = torch.zeros(n, device=DEV)
    for step in range(64):
        _lib.check(_lib.load().svit_dropout_mask(_lib.ptr(keep), n, 0.3, 1, step, 0xFFFF0000,
                                                _lib.vp(torch.cuda.current_stream().cuda_stream)), "mask")
        acc += keep.float() / 0.7
    assert abs(acc.mean().item() / 64 - 1.0) < 2e-3 and one._emb_drop_p == pytest.approx(0.3)


def test_dropout_mpp_matches_oracle_with_same_masks():
    """MPP forward / backward with dropout: emb_dropout at mpp.py:125, the encoder's dropouts through mpp.py:128."""
    from oracle.dropout import install
    from surface_vision_transformers_b200.mpp import draw_masks
    cfg = dict(dim=192, depth=3, heads=3, mlp_dim=768, num_patches=80, num_vertices=45)
    B, K = 4, 4 * 45
    torch.manual_seed(14)
    kw = dict(mask_prob=0.5, replace_prob=0.8, swap_prob=0.1, channels=4, num_vertices=cfg["num_vertices"])
    oracle = OracleMPP(OracleSiT(**cfg), cfg["dim"], K, DEV, **kw).to(DEV)
    for check_mode, tol in ((True, CHECK_TOL), (False, TOL)):
        ssl = svit.masked_patch_pretraining(transformer=svit.SiT(**cfg, dropout=0.1, emb_dropout=0.1), dim_in=cfg["dim"],
                                            dim_out=K, device=DEV, **kw)
        ssl.load_state_dict(oracle.state_dict())
        ssl.to(DEV)
        ssl.transformer.set_check_mode(check_mode).set_dropout_seed(31, 4)
        install(oracle.transformer, 0.1, 0.1, 31, 4)
        oracle.zero_grad()
        x = torch.randn(B, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
        masks = draw_masks(B, cfg["num_patches"], K, DEV, 0.5, 0.8, 0.1)
        lo, oo = oracle(x, masks=masks)
        lo.backward()
        lm, om = ssl(x, masks=masks)
        lm.backward()
        assert abs(lm.item() - lo.item()) / lo.item() < tol
        assert rel_l2(om, oo) < tol
        ref = dict(oracle.named_parameters())
        for n, p in ssl.named_parameters():
            if ref[n].grad is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            else:
                assert rel_l2(p.grad, ref[n].grad) < (tol if check_mode else 2 * tol), n
