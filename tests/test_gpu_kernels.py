"""-m gpu: kernel-level parity through the C ABI.  Integer / copy work is bit-exact; bf16 tensor-core work is
compared with fp32 torch math on the SAME bf16 operands (tolerance 1e-2 relative L2, north_star)."""
import hashlib

import numpy as np
import pytest
import torch

from helpers import load_golden, rel_l2

pytestmark = pytest.mark.gpu

import surface_vision_transformers_b200 as svit  # noqa: E402
from surface_vision_transformers_b200 import _lib  # noqa: E402
from surface_vision_transformers_b200._lib import check, ptr, vp  # noqa: E402

TOL = 1e-2


@pytest.fixture(scope="module")
def env():
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    return dict(dev=dev, lib=_lib.load(), sms=torch.cuda.get_device_properties(0).multi_processor_count)


def stream():
    return vp(torch.cuda.current_stream().cuda_stream)


def dgelu(x):
    return 0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * 0.3989422804014327 * torch.exp(-0.5 * x * x)


@pytest.mark.parametrize("M,N,K,mode,f32,bias,rowtab", [
    (128, 192, 64, 0, True, False, False),
    (128, 192, 64, 0, False, False, False),
    (1284, 1152, 384, 0, False, False, False),     # QKV projection (no bias)
    (1284, 384, 384, 2, True, True, False),        # to_out + bias + residual
    (1284, 1536, 384, 1, False, True, False),      # fc1 + bias + GELU (two outputs)
    (1284, 1536, 384, 4, False, True, False),      # fc1 + bias + GELU (inference)
    (1284, 384, 1536, 2, True, True, False),       # fc2 + bias + residual
    (1284, 1536, 384, 3, False, False, False),     # dgrad * gelu'
    (1284, 1536, 384, 5, False, True, False),      # fc1 + bias -> gelu' and gelu (what training stores)
    (20544, 1536, 384, 5, False, True, False),
    (1284, 1536, 384, 6, False, False, False),     # dgrad * stored gelu'
    (20544, 1536, 384, 6, False, False, False),
    (1284, 384, 640, 0, True, False, True),        # patch embedding + cls/pos table
    (5136, 612, 384, 0, True, True, False),        # MPP decoder (N not a tile multiple)
    (20544, 1152, 384, 0, False, False, False),    # several persistent waves (CTA pairs, A-resident schedule)
    (20544, 1536, 384, 1, False, True, False),     # fc1 + GELU at pair-tile size (A-resident only with SVIT_GEMM_ARES=1)
    (20544, 1536, 384, 3, False, False, False),    # dgrad * gelu' at pair-tile size
    (20544, 1536, 384, 4, False, True, False),     # GELU only at pair-tile size
    (20544, 612, 384, 0, True, True, False),       # A-resident: fp32 out, N not a tile multiple
    (19999, 1000, 328, 0, False, True, False),     # A-resident: ragged M, N and K (K not a multiple of 64)
    (20544, 768, 192, 1, False, True, False),      # A-resident: SiT-tiny fc1 (3 K blocks)
    (100, 100, 72, 0, True, True, False),          # ragged M, N, K
    (1, 192, 64, 0, True, True, False),            # single row
])
def test_gemm_tn(env, M, N, K, mode, f32, bias, rowtab):
    dev, lib = env["dev"], env["lib"]
    torch.manual_seed(M + N + K + mode)
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    odt = torch.float32 if f32 else torch.bfloat16
    out = torch.full((M, N), float("nan"), device=dev, dtype=odt)
    out2 = torch.full((M, N), float("nan"), device=dev, dtype=odt) if mode in (1, 5) else None
    aux = torch.randn(M, N, device=dev).to(odt) if mode in (2, 3, 6) else None
    period = 7
    rt = torch.randn(period, N, device=dev) if rowtab else None
    check(lib.svit_gemm_tn(ptr(A), ptr(B), ptr(out), ptr(out2), ptr(aux), ptr(b), ptr(rt), period, M, N, K, K, K, N, mode,
                           int(f32), env["sms"], stream()), "gemm_tn")
    torch.cuda.synchronize()
    acc = A.float() @ B.float().t()
    if bias:
        acc = acc + b
    if mode == 0:
        ref = acc + (rt[torch.arange(M, device=dev) % period] if rowtab else 0)
    elif mode == 1:
        assert rel_l2(out2, torch.nn.functional.gelu(acc.bfloat16().float())) < TOL
        ref = acc
    elif mode == 2:
        ref = acc + aux.float()
    elif mode == 3:
        ref = acc * dgelu(aux.float())
    elif mode == 5:
        assert rel_l2(out2, torch.nn.functional.gelu(acc)) < TOL
        ref = dgelu(acc)
    elif mode == 6:
        ref = acc * aux.float()
    else:
        ref = torch.nn.functional.gelu(acc)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < (1e-5 if (f32 and mode in (0, 2)) else TOL)


@pytest.mark.parametrize("M,N,K", [(64, 128, 192), (128, 128, 192), (1284, 1152, 384), (20544, 384, 1536),
                                   (1000, 104, 72), (5136, 384, 640), (3, 128, 64)])
def test_gemm_wgrad(env, M, N, K):
    dev, lib = env["dev"], env["lib"]
    torch.manual_seed(M + N + K)
    dY = (torch.randn(M, N, device=dev) * 0.5).bfloat16()
    X = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    dW = torch.zeros(N, K, device=dev)
    check(lib.svit_gemm_wgrad(ptr(dY), ptr(X), ptr(dW), M, N, K, N, K, K, env["sms"], stream()), "wgrad")
    torch.cuda.synchronize()
    assert rel_l2(dW, dY.float().t() @ X.float()) < 1e-4      # fp32 accumulation of exact bf16 products


@pytest.mark.parametrize("M,N,K", [(1284, 1536, 384), (700, 612, 384), (2000, 192, 192), (321, 128, 640), (64, 16, 8)])
def test_gemm_wgrad_bias(env, M, N, K):
    """dW and the fused bias gradient (column sums of dY via the ones-tile MMA); ragged pitches (ldy > N)."""
    dev, lib = env["dev"], env["lib"]
    torch.manual_seed(M + N + K + 1)
    ldy = (N + 15) // 8 * 8
    dYp = torch.zeros(M, ldy, device=dev, dtype=torch.bfloat16)
    dYp[:, :N] = (torch.randn(M, N, device=dev) * 0.5).bfloat16()
    dYp[:, N:] = 7.0   # padding columns must never be read into the result
    X = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    dW = torch.zeros(N, K, device=dev)
    db = torch.zeros(N, device=dev)
    check(lib.svit_gemm_wgrad_bias(ptr(dYp), ptr(X), ptr(dW), ptr(db), M, N, K, ldy, K, K, env["sms"], stream()), "wgrad_bias")
    torch.cuda.synchronize()
    dY = dYp[:, :N].float()
    assert rel_l2(dW, dY.t() @ X.float()) < 1e-4
    assert rel_l2(db, dY.sum(0)) < 1e-4


def test_gemm_rejects_bad_pitch(env):
    dev, lib = env["dev"], env["lib"]
    A = torch.zeros(8, 100, device=dev, dtype=torch.bfloat16)
    rc = lib.svit_gemm_tn(ptr(A), ptr(A), ptr(A), vp(0), vp(0), vp(0), vp(0), 1, 8, 8, 100, 100, 100, 100, 0, 0, 148, stream())
    assert rc != 0 and b"16-byte" in lib.svit_last_error()


@pytest.mark.parametrize("B,H,T", [(1, 1, 128), (2, 3, 81), (2, 2, 321), (3, 6, 321), (1, 1, 21), (2, 2, 384), (1, 2, 200), (2, 1, 1)])
def test_attention_fwd_bwd(env, B, H, T):
    dev, lib = env["dev"], env["lib"]
    torch.manual_seed(B * 1000 + H * 10 + T)
    inner = H * 64
    scale = 64 ** -0.5
    qkv = torch.randn(B, T, 3 * inner, device=dev).bfloat16()
    out = torch.full((B, T, inner), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev)
    check(lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, scale, stream()), "attn_fwd")
    q, k, v = [t.reshape(B, T, H, 64).permute(0, 2, 1, 3).float().requires_grad_(True) for t in qkv.chunk(3, dim=-1)]
    dots = q @ k.transpose(-1, -2) * scale
    ref = (dots.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, T, inner)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < TOL
    assert rel_l2(lse, torch.logsumexp(dots, -1)) < 1e-3
    dout = torch.randn(B, T, inner, device=dev).bfloat16()
    dqkv = torch.full((B, T, 3 * inner), float("nan"), device=dev, dtype=torch.bfloat16)
    check(lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, scale, stream()), "attn_bwd")
    ref.backward(dout.float())
    torch.cuda.synchronize()
    dref = torch.cat([g.permute(0, 2, 1, 3).reshape(B, T, inner) for g in (q.grad, k.grad, v.grad)], dim=-1)
    assert torch.isfinite(dqkv.float()).all()
    for i in range(3):
        got, want = dqkv[..., i * inner:(i + 1) * inner].float(), dref[..., i * inner:(i + 1) * inner]
        if T == 1 and i < 2:
            assert got.abs().max() < 1e-5            # softmax over one key: dq = dk = 0 exactly in the reference
        else:
            assert rel_l2(got, want) < 2 * TOL


@pytest.mark.parametrize("B,H,T", [(1, 1, 1), (2, 3, 21), (3, 6, 321), (2, 12, 321), (2, 2, 384), (1, 2, 257)])
def test_attention_cls_row_fwd_bwd(env, B, H, T):
    """svit_attn_cls_fwd / _bwd (the last block under cls pooling): attention of the query row of token 0 against every
    key, and its backward -- dk, dv of all keys, dq of token 0, exact zeros in the other dq rows -- against fp32 torch on
    the same bf16 inputs."""
    dev, lib = env["dev"], env["lib"]
    torch.manual_seed(B * 1000 + H * 10 + T)
    inner = H * 64
    scale = 64 ** -0.5
    qkv = torch.randn(B, T, 3 * inner, device=dev).bfloat16()
    out = torch.full((B, inner), float("nan"), device=dev, dtype=torch.bfloat16)
    prob = torch.full((B, H, T), float("nan"), device=dev)
    check(lib.svit_attn_cls_fwd(ptr(qkv), ptr(out), ptr(prob), B, H, T, scale, stream()), "attn_cls_fwd")
    q, k, v = [t.reshape(B, T, H, 64).permute(0, 2, 1, 3).float().requires_grad_(True) for t in qkv.chunk(3, dim=-1)]
    p_ref = (q[:, :, :1] @ k.transpose(-1, -2) * scale).softmax(-1)            # (B, H, 1, T)
    ref = (p_ref @ v).permute(0, 2, 1, 3).reshape(B, inner)
    torch.cuda.synchronize()
    assert rel_l2(prob, p_ref.squeeze(2)) < 1e-4 and rel_l2(out, ref) < 4e-3
    dout = torch.randn(B, inner, device=dev).bfloat16()
    dqkv = torch.full((B, T, 3 * inner), float("nan"), device=dev, dtype=torch.bfloat16)
    check(lib.svit_attn_cls_bwd(ptr(qkv), ptr(prob), ptr(dout), ptr(dqkv), B, H, T, scale, stream()), "attn_cls_bwd")
    ref.backward(dout.float())
    torch.cuda.synchronize()
    dref = torch.cat([g.permute(0, 2, 1, 3).reshape(B, T, inner) for g in (q.grad, k.grad, v.grad)], dim=-1)
    assert torch.isfinite(dqkv.float()).all()
    assert (dqkv[:, 1:, :inner] == 0).all()                                    # rows that were never queries
    if T > 1:
        for i in range(3):
            assert rel_l2(dqkv[..., i * inner:(i + 1) * inner].float(), dref[..., i * inner:(i + 1) * inner]) < 5e-3, i
    else:
        assert dqkv[..., :2 * inner].float().abs().max() < 1e-5 and rel_l2(dqkv[..., 2 * inner:].float(), dref[..., 2 * inner:]) < 5e-3


@pytest.mark.parametrize("T", [2, 31, 33, 95, 96, 97, 112, 127, 129, 191, 192, 193, 255, 256, 257, 287, 288, 289, 320, 322, 383])
def test_attention_boundary_lengths(env, T):
    """Every tile boundary of the fused kernels (96-key blocks x 32-column slabs and 128-query blocks in the backward,
    32-column groups in the forward): one token short of, exactly at and one past each of them.  NaN-prefilled outputs
    catch rows or columns that a masked path forgot; a poisoned second pass catches stale TMEM contents."""
    dev, lib = env["dev"], env["lib"]
    B, H = 2, 2
    torch.manual_seed(T)
    inner, scale = H * 64, 64 ** -0.5
    qkv = torch.randn(B, T, 3 * inner, device=dev).bfloat16()
    dout = torch.randn(B, T, inner, device=dev).bfloat16()
    q, k, v = [t.reshape(B, T, H, 64).permute(0, 2, 1, 3).float().requires_grad_(True) for t in qkv.chunk(3, dim=-1)]
    dots = q @ k.transpose(-1, -2) * scale
    ref = (dots.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, T, inner)
    ref.backward(dout.float())
    dref = torch.cat([g.permute(0, 2, 1, 3).reshape(B, T, inner) for g in (q.grad, k.grad, v.grad)], dim=-1)
    # a first launch on a long, huge-valued problem leaves large numbers in every TMEM column / smem tile
    big = (torch.randn(B, 384, 3 * inner, device=dev) * 30).bfloat16()
    bo = torch.empty(B, 384, inner, device=dev, dtype=torch.bfloat16)
    bl = torch.zeros(B, H, 384, device=dev)
    bd = torch.empty_like(big)
    for _ in range(2):
        check(lib.svit_attn_fwd(ptr(big), ptr(bo), ptr(bl), B, H, 384, scale, stream()), "attn_fwd")
        check(lib.svit_attn_bwd(ptr(big), ptr(bo), ptr(bo), ptr(bl), ptr(bd), B, H, 384, scale, stream()), "attn_bwd")
        out = torch.full((B, T, inner), float("nan"), device=dev, dtype=torch.bfloat16)
        lse = torch.full((B, H, T), float("nan"), device=dev)
        dqkv = torch.full((B, T, 3 * inner), float("nan"), device=dev, dtype=torch.bfloat16)
        check(lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, scale, stream()), "attn_fwd")
        check(lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, scale, stream()), "attn_bwd")
        torch.cuda.synchronize()
        assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all() and torch.isfinite(dqkv.float()).all()
        assert rel_l2(out, ref) < TOL
        assert rel_l2(lse, torch.logsumexp(dots, -1)) < 1e-3
        for i in range(3):
            assert rel_l2(dqkv[..., i * inner:(i + 1) * inner].float(), dref[..., i * inner:(i + 1) * inner]) < 2 * TOL, i


def test_attention_bitwise_repeatable_at_full_size(env):
    """Neither attention kernel uses atomics, so repeated launches on the same inputs must agree bit for bit -- a cheap
    race detector for the hand-rolled TMEM / shared-memory hand-offs (a missed barrier shows up as run-to-run noise).
    BASELINE.json configs[1] shape per GPU (B = 256, H = 6, T = 321); a different problem runs in between so that TMEM,
    shared memory and L2 start from different contents each time."""
    dev, lib = env["dev"], env["lib"]
    B, H, T = 256, 6, 321
    torch.manual_seed(5)
    inner, scale = H * 64, 64 ** -0.5
    qkv = torch.randn(B, T, 3 * inner, device=dev).bfloat16()
    dout = torch.randn(B, T, inner, device=dev).bfloat16()
    other = (torch.randn(B, 200, 3 * inner, device=dev) * 3).bfloat16()
    oo = torch.empty(B, 200, inner, device=dev, dtype=torch.bfloat16)
    ol = torch.empty(B, H, 200, device=dev)
    od = torch.empty_like(other)
    ref = None
    for it in range(4):
        out = torch.empty(B, T, inner, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, H, T, device=dev)
        dqkv = torch.empty_like(qkv)
        check(lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, scale, stream()), "attn_fwd")
        check(lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, scale, stream()), "attn_bwd")
        check(lib.svit_attn_fwd(ptr(other), ptr(oo), ptr(ol), B, H, 200, scale, stream()), "attn_fwd")
        check(lib.svit_attn_bwd(ptr(other), ptr(oo), ptr(oo), ptr(ol), ptr(od), B, H, 200, scale, stream()), "attn_bwd")
        torch.cuda.synchronize()
        cur = (out.clone(), lse.clone(), dqkv.clone())
        if ref is None:
            ref = cur
        else:
            for a, b_ in zip(cur, ref):
                assert torch.equal(a, b_), it
    assert torch.isfinite(ref[2].float()).all()


@pytest.mark.parametrize("kind", ["large_random", "ascending_keys", "descending_keys"])
def test_attention_fwd_extreme_logits(env, kind):
    """The online softmax with lazy rescaling stays exact when later key chunks raise the row maximum by far more than
    2^8 (ascending keys force the rescale of O in every chunk), when the first chunk dominates, and for very peaked
    random rows.  Output and log-sum-exp against the fp32 reference on the same bf16 inputs."""
    dev, lib = env["dev"], env["lib"]
    torch.manual_seed(7)
    B, H, T = 2, 2, 321
    inner = H * 64
    scale = 64 ** -0.5
    if kind == "large_random":
        qkv = torch.randn(B, T, 3 * inner, device=dev) * 6.0
    else:
        u = torch.nn.functional.normalize(torch.randn(64, device=dev), dim=0)
        ramp = torch.linspace(0.0, 1.0, T, device=dev)
        if kind == "descending_keys":
            ramp = ramp.flip(0)
        q = (u * 40.0).expand(B, T, H, 64) + torch.randn(B, T, H, 64, device=dev)
        k = (u * 60.0)[None, None, None, :] * ramp[None, :, None, None] + torch.randn(B, T, H, 64, device=dev)
        v = torch.randn(B, T, H, 64, device=dev)
        qkv = torch.cat([t.reshape(B, T, inner) for t in (q, k, v)], dim=-1)
    qkv = qkv.bfloat16()
    out = torch.full((B, T, inner), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev)
    check(lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, scale, stream()), "attn_fwd")
    q, k, v = [t.reshape(B, T, H, 64).permute(0, 2, 1, 3).float() for t in qkv.chunk(3, dim=-1)]
    dots = q @ k.transpose(-1, -2) * scale
    ref = (dots.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, T, inner)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert rel_l2(out, ref) < TOL
    assert rel_l2(lse, torch.logsumexp(dots, -1)) < 1e-3


def test_attention_rejects_long_sequences(env):
    lib = env["lib"]
    assert lib.svit_attn_fwd(vp(256), vp(256), vp(256), 1, 1, 385, 0.125, stream()) != 0
    assert b"unsupported shape" in lib.svit_last_error()


@pytest.mark.parametrize("M,D", [(321 * 3, 384), (77, 192), (129, 768), (5, 128)])
def test_layernorm_fwd_bwd(env, M, D):
    dev, lib = env["dev"], env["lib"]
    torch.manual_seed(M + D)
    x = (torch.randn(M, D, device=dev) * 2 + 0.5).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(D, device=dev)).requires_grad_(True)
    beta = (0.1 * torch.randn(D, device=dev)).requires_grad_(True)
    a = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    check(lib.svit_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(a), ptr(mean), ptr(rstd), M, D, 1e-5, stream()), "ln_fwd")
    ref = torch.nn.functional.layer_norm(x, (D,), gamma, beta, 1e-5)
    torch.cuda.synchronize()
    assert rel_l2(a, ref) < 4e-3                           # bf16 rounding of the output only
    assert rel_l2(mean, x.mean(1)) < 1e-5
    assert rel_l2(rstd, (x.var(1, unbiased=False) + 1e-5).rsqrt()) < 1e-5
    da = torch.randn(M, D, device=dev).bfloat16()
    g_in = torch.randn(M, D, device=dev)
    g_out = torch.empty(M, D, device=dev); g16 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); cs = torch.zeros(D, device=dev)
    check(lib.svit_layernorm_bwd(ptr(da), ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(g_in), ptr(g_out), ptr(g16), ptr(dg),
                                 ptr(db), ptr(cs), M, D, stream()), "ln_bwd")
    ref.backward(da.float())
    torch.cuda.synchronize()
    assert rel_l2(g_out, g_in + x.grad) < 1e-4
    assert rel_l2(g16, g_in + x.grad) < 4e-3
    assert rel_l2(dg, gamma.grad) < 1e-4 and rel_l2(db, beta.grad) < 1e-4
    assert rel_l2(cs, (g_in + x.grad).sum(0)) < 1e-4


@pytest.mark.parametrize("sub_ico", [1, 2])
def test_gather_bit_exact_vs_golden(env, sub_ico):
    """a1: on-device patch gather == the reference's preprocessing loop, bit for bit (SHA-256 of the fp32 result)."""
    dev = env["dev"]
    g = load_golden("gather")
    table = svit.load_index_table(sub_ico, dev)
    rs = np.random.RandomState(100 + sub_ico)
    data = rs.standard_normal((4, 4, 40962)).astype(np.float32)
    means = np.array([1.15, 0.037, 1.0, 0.07], dtype=np.float32).reshape(1, 4, 1)
    stds = np.array([0.41, 0.19, 0.39, 4.05], dtype=np.float32).reshape(1, 4, 1)
    normalised = ((data - means) / stds).astype(np.float32)
    out = svit.gather_patches(torch.from_numpy(normalised).to(dev), table)
    from surface_vision_transformers_b200.gather import preprocessing_layout
    res = preprocessing_layout(out).cpu().numpy()
    assert res.shape == tuple(g[f"shape/{sub_ico}"])
    assert hashlib.sha256(res.tobytes()).hexdigest() == str(g[f"sha256_f32/{sub_ico}"])


def test_gather_empty_and_oracle(env):
    from oracle import gather_oracle
    dev = env["dev"]
    table = svit.load_index_table(2, dev)
    assert svit.gather_patches(torch.zeros(0, 4, 40962, device=dev), table).shape == (0, 4, 320, 153)
    mesh = torch.randn(3, 4, 40962, device=dev)
    out = svit.gather_patches(mesh, table)
    ref = gather_oracle.gather_patches(mesh.cpu().numpy(), table.cpu().numpy().astype(np.int64))
    assert np.array_equal(out.cpu().numpy(), ref)


def test_adamw_kernel_matches_torch(env):
    dev = env["dev"]
    torch.manual_seed(0)
    cfg = dict(dim=128, depth=2, heads=2, mlp_dim=128, num_patches=6, num_vertices=5)
    m = svit.SiT(**cfg).to(dev)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in m.parameters()]
    a = svit.FusedAdamW(m.parameters(), lr=1e-2, weight_decay=0.05)
    b = torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.05)
    for it in range(3):
        for i, (p, q) in enumerate(zip(m.parameters(), ref)):
            if it == 1 and i >= len(ref) - 4:          # params without gradient are skipped (MPP: mlp_head)
                p.grad = None; q.grad = None
                continue
            g = torch.randn_like(q)
            p.grad = g.clone(); q.grad = g.clone()
        a.step(); b.step()
    torch.cuda.synchronize()
    for p, q in zip(m.parameters(), ref):
        assert torch.allclose(p, q, rtol=2e-5, atol=1e-6)
    sd = a.state_dict()
    assert len(sd["state"]) == len(ref)


@pytest.mark.parametrize("M,K", [(845, 384), (845, 1536), (5, 384), (256, 64), (20544, 1536), (82176, 384)])
def test_gemm_ln_residual_linear_plus_layernorm(env, M, K):
    """svit_gemm_ln (gemm_ln.cu): x_out = A W^T + bias + x_in in fp32 and, from the same accumulator tile, the LayerNorm
    of the next sub-layer (bf16 output, mean, rstd) -- against fp32 torch on the same bf16 operands.  Ragged M (partial
    256-row tile), a single K block, several persistent waves."""
    dev, lib = env["dev"], env["lib"]
    D = 384
    torch.manual_seed(M + K)
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    W = (torch.randn(D, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(D, device=dev) * 0.1
    x_in = torch.randn(M, D, device=dev) * 2 + 0.3
    gamma = torch.rand(D, device=dev) + 0.5
    beta = torch.randn(D, device=dev) * 0.1
    x_out = torch.full((M, D), float("nan"), device=dev)
    a_out = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    mean = torch.empty(M, device=dev)
    rstd = torch.empty(M, device=dev)
    check(lib.svit_gemm_ln(ptr(A), ptr(W), ptr(bias), ptr(x_in), ptr(x_out), ptr(a_out), ptr(gamma), ptr(beta), ptr(mean),
                           ptr(rstd), M, D, K, K, K, 1e-5, env["sms"], stream()), "gemm_ln")
    torch.cuda.synchronize()
    ref_x = A.float() @ W.float().t() + bias + x_in
    assert rel_l2(x_out, ref_x) < 1e-5
    ref_mean = ref_x.mean(1)
    ref_rstd = (ref_x.var(1, unbiased=False) + 1e-5).rsqrt()
    assert rel_l2(mean, ref_mean) < 1e-5 and rel_l2(rstd, ref_rstd) < 1e-5
    ref_a = torch.nn.functional.layer_norm(ref_x, (D,), gamma, beta, 1e-5)
    assert rel_l2(a_out.float(), ref_a) < 4e-3        # bf16 rounding of the output only
    # the stand-alone kernels give the same answer (what the fused launch replaces)
    x2 = torch.empty(M, D, device=dev)
    a2 = torch.empty_like(a_out); m2 = torch.empty(M, device=dev); r2 = torch.empty(M, device=dev)
    check(lib.svit_gemm_tn(ptr(A), ptr(W), ptr(x2), vp(0), ptr(x_in), ptr(bias), vp(0), 1, M, D, K, K, K, D, 2, 1, env["sms"],
                           stream()), "gemm_tn")
    check(lib.svit_layernorm_fwd(ptr(x2), ptr(gamma), ptr(beta), ptr(a2), ptr(m2), ptr(r2), M, D, 1e-5, stream()), "ln")
    torch.cuda.synchronize()
    assert rel_l2(x_out, x2) < 1e-6 and rel_l2(a_out.float(), a2.float()) < 2e-3
