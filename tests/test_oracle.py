"""CPU tests that pin the oracle: against the reference's own modules (when /root/reference is present), against
the golden vectors generated from them, and -- for the third-party encoder the reference does not ship -- against
independent implementations (HuggingFace ViT layers, torch SDPA)."""
import hashlib
import math
import os

import numpy as np
import pytest
import torch

from helpers import CASES, load_golden, rel_l2, seeded_input, seeded_masks, seeded_state, subsample
from oracle import gather_oracle, reference_loader, vit_shim
from oracle.sit_oracle import OracleMPP, OracleSiT, draw_mpp_masks

needs_reference = pytest.mark.skipif(not reference_loader.available(), reason="/root/reference not present")


def build_oracle(case):
    cfg = case["cfg"]
    torch.manual_seed(0)
    model = OracleSiT(**cfg)
    if case["kind"] == "sit":
        model.load_state_dict(seeded_state(model, case["seed"]))
        return model, None
    K = cfg["num_channels"] * cfg["num_vertices"]
    ssl = OracleMPP(model, cfg["dim"], K, "cpu", channels=cfg["num_channels"], num_vertices=cfg["num_vertices"],
                    **case["mpp"])
    ssl.load_state_dict(seeded_state(ssl, case["seed"]))
    return model, ssl


@pytest.mark.parametrize("name", ["sit_cls", "sit_mean"])
def test_oracle_sit_matches_golden(name):
    case = CASES[name]
    g = load_golden(name)
    model, _ = build_oracle(case)
    x, y = seeded_input(case["cfg"], case["batch"], case["seed"])
    pred = model(x)
    loss = torch.nn.functional.mse_loss(pred.squeeze(), y)
    loss.backward()
    assert rel_l2(pred.detach(), g["pred"]) < 1e-5
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < 1e-5
    for k, p in model.named_parameters():
        assert bool(g["grad_none/" + k]) == (p.grad is None)
        assert rel_l2(subsample(p.grad), g["grad_sub/" + k]) < 1e-4, k
        assert abs(p.grad.norm().item() - float(g["grad_norm/" + k])) <= 1e-4 * float(g["grad_norm/" + k]) + 1e-7, k


def test_oracle_mpp_matches_golden():
    case = CASES["mpp"]
    g = load_golden("mpp")
    _, ssl = build_oracle(case)
    x, _ = seeded_input(case["cfg"], case["batch"], case["seed"])
    masks = tuple(torch.from_numpy(g[k]) for k in ("mask", "swap_sel", "swap_src", "replace_sel"))
    loss, out = ssl(x, masks=masks)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < 1e-5
    assert rel_l2(out.detach(), g["batch_out"]) < 1e-5
    for k, p in ssl.named_parameters():
        assert bool(g["grad_none/" + k]) == (p.grad is None), k
        if p.grad is not None:
            assert rel_l2(subsample(p.grad), g["grad_sub/" + k]) < 1e-4, k
    # seeded RNG path: same torch RNG call order as models/mpp.py -> same loss as the reference's own forward
    _, ssl2 = build_oracle(case)
    torch.manual_seed(case["seed"])
    loss_rng, out_rng = ssl2(x)
    assert abs(loss_rng.item() - float(g["loss_rng"])) / float(g["loss_rng"]) < 1e-5
    assert rel_l2(subsample(out_rng), g["batch_out_rng_sub"]) < 1e-5


@needs_reference
@pytest.mark.parametrize("name", ["sit_cls", "sit_mean"])
def test_oracle_sit_equals_reference_module(name):
    SiT, _, _ = reference_loader.load_reference_models()
    case = CASES[name]
    ref = SiT(**case["cfg"])
    orc = OracleSiT(**case["cfg"])
    assert list(ref.state_dict().keys()) == list(orc.state_dict().keys())
    sd = seeded_state(ref, case["seed"])
    ref.load_state_dict(sd)
    orc.load_state_dict(sd)
    x, _ = seeded_input(case["cfg"], case["batch"], case["seed"])
    assert torch.equal(ref(x), orc(x))


@needs_reference
def test_oracle_mpp_equals_reference_module_same_rng():
    SiT, MPP, _ = reference_loader.load_reference_models()
    case = CASES["mpp"]
    cfg = case["cfg"]
    K = cfg["num_channels"] * cfg["num_vertices"]
    ref = MPP(transformer=SiT(**cfg), dim_in=cfg["dim"], dim_out=K, device="cpu", channels=cfg["num_channels"],
              num_vertices=cfg["num_vertices"], **case["mpp"])
    orc = OracleMPP(OracleSiT(**cfg), cfg["dim"], K, "cpu", channels=cfg["num_channels"],
                    num_vertices=cfg["num_vertices"], **case["mpp"])
    assert list(ref.state_dict().keys()) == list(orc.state_dict().keys())
    sd = seeded_state(ref, case["seed"])
    ref.load_state_dict(sd)
    orc.load_state_dict(sd)
    x, _ = seeded_input(cfg, case["batch"], case["seed"])
    torch.manual_seed(5)
    l1, o1 = ref(x)
    torch.manual_seed(5)
    l2, o2 = orc(x)
    assert torch.equal(l1, l2) and torch.equal(o1, o2)
    l1.backward()
    l2.backward()
    for (k, p), (_, q) in zip(ref.named_parameters(), orc.named_parameters()):
        assert (p.grad is None) == (q.grad is None), k
        if p.grad is not None:
            assert torch.allclose(p.grad, q.grad, rtol=1e-6, atol=1e-8), k


def test_state_dict_keys_match_utils_mapping():
    """Key patterns written by the reference's ImageNet remapper (utils/utils.py:13-33) exist with the right shapes."""
    depth, dim, mlp = 3, 128, 256
    m = OracleSiT(dim=dim, depth=depth, heads=2, mlp_dim=mlp, num_patches=5, num_vertices=6)
    sd = m.state_dict()
    assert sd["mlp_head.0.weight"].shape == (dim,) and sd["mlp_head.0.bias"].shape == (dim,)
    for i in range(depth):
        p = f"transformer.layers.{i}."
        assert sd[p + "0.norm.weight"].shape == (dim,) and sd[p + "1.norm.bias"].shape == (dim,)
        assert sd[p + "0.fn.to_qkv.weight"].shape == (3 * 2 * 64, dim)
        assert p + "0.fn.to_qkv.bias" not in sd
        assert sd[p + "0.fn.to_out.0.weight"].shape == (dim, 2 * 64) and sd[p + "0.fn.to_out.0.bias"].shape == (dim,)
        assert sd[p + "1.fn.net.0.weight"].shape == (mlp, dim) and sd[p + "1.fn.net.3.weight"].shape == (dim, mlp)
    assert len(sd) == 8 + 11 * depth
    assert not any(k.startswith("transformer.norm") for k in sd)


def test_vit_shim_matches_huggingface_vit_layers():
    """Independent pin of the absent third-party encoder: HF ViT layers (pre-norm, exact GELU, no qkv bias) with the
    same weights must give the same output -- the equivalence utils/utils.py::load_weights_imagenet relies on."""
    transformers = pytest.importorskip("transformers")
    from transformers.models.vit.configuration_vit import ViTConfig
    from transformers.models.vit.modeling_vit import ViTLayer
    dim, depth, heads, mlp = 128, 2, 2, 256
    torch.manual_seed(3)
    shim = vit_shim.Transformer(dim, depth, heads, 64, mlp).eval()
    cfg = ViTConfig(hidden_size=dim, num_hidden_layers=depth, num_attention_heads=heads, intermediate_size=mlp,
                    hidden_act="gelu", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, qkv_bias=False,
                    layer_norm_eps=1e-5)
    cfg._attn_implementation = "eager"
    x = torch.randn(2, 9, dim)
    y = x
    for (attn, ff) in shim.layers:
        layer = ViTLayer(cfg).eval()
        sd = layer.state_dict()
        wq, wk, wv = attn.fn.to_qkv.weight.chunk(3, dim=0)
        new = {
            "layernorm_before.weight": attn.norm.weight, "layernorm_before.bias": attn.norm.bias,
            "attention.attention.query.weight": wq, "attention.attention.key.weight": wk,
            "attention.attention.value.weight": wv,
            "attention.output.dense.weight": attn.fn.to_out[0].weight, "attention.output.dense.bias": attn.fn.to_out[0].bias,
            "layernorm_after.weight": ff.norm.weight, "layernorm_after.bias": ff.norm.bias,
            "intermediate.dense.weight": ff.fn.net[0].weight, "intermediate.dense.bias": ff.fn.net[0].bias,
            "output.dense.weight": ff.fn.net[3].weight, "output.dense.bias": ff.fn.net[3].bias,
        }
        assert set(new) == set(sd), (set(sd) ^ set(new))
        layer.load_state_dict({k: v.detach().clone() for k, v in new.items()})
        out = layer(y)
        y = out[0] if isinstance(out, tuple) else out
    ref = shim(x)
    assert torch.allclose(ref, y, rtol=1e-5, atol=1e-5), (ref - y).abs().max()


def test_vit_shim_attention_matches_sdpa():
    torch.manual_seed(4)
    att = vit_shim.Attention(128, heads=2, dim_head=64)
    x = torch.randn(3, 21, 128)
    qkv = att.to_qkv(x).chunk(3, dim=-1)
    q, k, v = (t.reshape(3, 21, 2, 64).transpose(1, 2) for t in qkv)
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v)   # default scale = 64 ** -0.5
    o = att.to_out(o.transpose(1, 2).reshape(3, 21, 128))
    assert torch.allclose(att(x), o, rtol=1e-5, atol=1e-6)


def test_mask_semantics():
    """mpp.py:25-39: exactly ceil(p*N) masked tokens per sample; replace wins over swap (Appendix B)."""
    torch.manual_seed(0)
    b, n, k = 5, 13, 8
    like = torch.zeros(b, n, k)
    mask, swap_sel, swap_src, replace_sel = draw_mpp_masks(like, 0.5, 0.8, 0.02)
    assert mask.dtype == torch.bool and mask.shape == (b, n)
    assert (mask.sum(1) == math.ceil(0.5 * n)).all()
    assert not (swap_sel & ~mask).any() and not (replace_sel & ~mask).any()
    assert swap_src.min() >= 0 and swap_src.max() < n


# ------------------------------------------------------------------------------------------- gather
def _table(sub_ico):
    from surface_vision_transformers_b200.gather import load_index_table
    return load_index_table(sub_ico).numpy().astype(np.int64)


@pytest.mark.parametrize("sub_ico,shape,dups", [(1, (561, 80), 3918), (2, (153, 320), 7998)])
def test_index_tables_structure(sub_ico, shape, dups):
    """Structural pins of the gather tables recorded in SURVEY.md section 4."""
    t = _table(sub_ico)
    assert t.shape == shape
    assert t.min() == 0 and t.max() == 40961
    assert len(np.unique(t)) == 40962                      # every ico-6 vertex is covered
    for j in range(t.shape[1]):
        assert len(np.unique(t[:, j])) == t.shape[0]       # no duplicate inside a patch
    assert t.size - 40962 == dups


@needs_reference
@pytest.mark.parametrize("sub_ico", [1, 2])
def test_index_tables_equal_reference_csv(sub_ico):
    import pandas as pd
    df = pd.read_csv(os.path.join(reference_loader.REFERENCE_ROOT, "utils", f"triangle_indices_ico_6_sub_ico_{sub_ico}.csv"))
    t = _table(sub_ico)
    for j in range(t.shape[1]):
        assert np.array_equal(df[str(j)].to_numpy(), t[:, j])


@pytest.mark.parametrize("sub_ico", [1, 2])
def test_gather_oracle_matches_golden(sub_ico):
    g = load_golden("gather")
    t = _table(sub_ico)
    rs = np.random.RandomState(100 + sub_ico)
    data = rs.standard_normal((4, 4, 40962)).astype(np.float32)
    means = np.array([1.15, 0.037, 1.0, 0.07], dtype=np.float32).reshape(1, 4, 1)
    stds = np.array([0.41, 0.19, 0.39, 4.05], dtype=np.float32).reshape(1, 4, 1)
    res = gather_oracle.preprocessing_layout(gather_oracle.zscore(data, means, stds), t).astype(np.float32)
    assert tuple(g[f"shape/{sub_ico}"]) == res.shape
    assert hashlib.sha256(res.tobytes()).hexdigest() == str(g[f"sha256_f32/{sub_ico}"])
    assert np.array_equal(res.reshape(-1)[::997], g[f"sub/{sub_ico}"])
