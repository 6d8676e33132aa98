"""Micro-benchmarks of individual kernels at the SiT-small ico-2 B=256 shapes (CUDA-event timed)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surface_vision_transformers_b200 import _lib
from surface_vision_transformers_b200._lib import ptr, vp, check
lib = _lib.load()
dev = torch.device("cuda:0")
SMS = torch.cuda.get_device_properties(0).multi_processor_count
B = int(os.environ.get("B", 256)); T = 321; D = 384; H = 6; MLP = 1536
M = B * T
def st(): return vp(torch.cuda.current_stream().cuda_stream)

def timeit(name, f, bytes_=None, flops=None, iters=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    extra = ""
    if bytes_: extra += f"  {bytes_/ms/1e6:8.1f} GB/s"
    if flops: extra += f"  {flops/ms/1e9:8.1f} TFLOP/s"
    print(f"{name:38s} {ms*1e3:9.1f} us{extra}", flush=True)

which = sys.argv[1:] or ["ln", "gemm", "wgrad", "attn", "misc"]
if "ln" in which:
    x = torch.randn(M, D, device=dev); g = torch.ones(D, device=dev); bt = torch.zeros(D, device=dev)
    a = torch.empty(M, D, device=dev, dtype=torch.bfloat16); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    timeit("ln_fwd", lambda: lib.svit_layernorm_fwd(ptr(x), ptr(g), ptr(bt), ptr(a), ptr(mean), ptr(rstd), M, D, 1e-5, st()), bytes_=M*D*6)
    da = torch.randn(M, D, device=dev).bfloat16(); gi = torch.randn(M, D, device=dev); go = torch.empty_like(gi); g16 = torch.empty_like(a)
    dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); cs = torch.zeros(D, device=dev)
    timeit("ln_bwd", lambda: lib.svit_layernorm_bwd(ptr(da), ptr(x), ptr(mean), ptr(rstd), ptr(g), ptr(gi), ptr(go), ptr(g16), ptr(dg), ptr(db), ptr(cs), M, D, st()), bytes_=M*D*(2+4+4+4+2))
if "gemm" in which:
    def gemm(name, N, K, mode, f32):
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        bias = torch.zeros(N, device=dev); odt = torch.float32 if f32 else torch.bfloat16
        o = torch.empty(M, N, device=dev, dtype=odt); o2 = torch.empty(M, N, device=dev, dtype=odt) if mode in (1, 5) else None
        aux = torch.randn(M, N, device=dev).to(odt) if mode in (2, 3, 6) else None
        osz = 4 if f32 else 2
        by = M*K*2 + N*K*2 + M*N*osz*(2 if mode in (1, 5) else 1) + (M*N*osz if aux is not None else 0)
        timeit(name, lambda: lib.svit_gemm_tn(ptr(A), ptr(W), ptr(o), ptr(o2), ptr(aux), ptr(bias), vp(0), 1, M, N, K, K, K, N, mode, int(f32), SMS, st()),
               bytes_=by, flops=2.0*M*N*K)
    gemm("gemm qkv   [M,384]x[1152] store bf16", 1152, 384, 0, False)
    gemm("gemm out   [M,384]x[384] resid f32", 384, 384, 2, True)
    gemm("gemm fc1   [M,384]x[1536] gelu x2", 1536, 384, 1, False)
    gemm("gemm fc1   [M,384]x[1536] gelu only", 1536, 384, 4, False)
    gemm("gemm fc2   [M,1536]x[384] resid f32", 384, 1536, 2, True)
    gemm("gemm fc1   [M,384]x[1536] gelu' + gelu", 1536, 384, 5, False)
    gemm("gemm dfc2  [M,384]x[1536] dgelu", 1536, 384, 3, False)
    gemm("gemm dfc2  [M,384]x[1536] * stored gelu'", 1536, 384, 6, False)
    gemm("gemm dfc1  [M,1536]x[384] store bf16", 384, 1536, 0, False)
    gemm("gemm dqkv  [M,1152]x[384] store bf16", 384, 1152, 0, False)
if "gemm" in which:
    # residual Linear + the next LayerNorm in one kernel (gemm_ln.cu) vs the two launches it replaces
    for (K, nm) in ((384, "out"), (1536, "fc2")):
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); W = (torch.randn(D, K, device=dev) * 0.05).bfloat16()
        bias = torch.zeros(D, device=dev); xin = torch.randn(M, D, device=dev); xo = torch.empty(M, D, device=dev)
        ao = torch.empty(M, D, device=dev, dtype=torch.bfloat16); g = torch.ones(D, device=dev); bt = torch.zeros(D, device=dev)
        mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
        by = M*K*2 + D*K*2 + M*D*4*2 + M*D*2
        timeit(f"gemm_ln {nm}  [M,{K}]x[384] resid f32 + LN", lambda: lib.svit_gemm_ln(ptr(A), ptr(W), ptr(bias), ptr(xin), ptr(xo), ptr(ao), ptr(g), ptr(bt), ptr(mean), ptr(rstd), M, D, K, K, K, 1e-5, SMS, st()),
               bytes_=by, flops=2.0*M*D*K)
if "wgrad" in which:
    for (N, K) in [(1152, 384), (384, 384), (1536, 384), (384, 1536), (384, 640)]:
        dY = (torch.randn(M, N, device=dev) * 0.5).bfloat16(); X = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        dW = torch.zeros(N, K, device=dev)
        timeit(f"wgrad dW[{N},{K}]", lambda: lib.svit_gemm_wgrad(ptr(dY), ptr(X), ptr(dW), M, N, K, N, K, K, SMS, st()), bytes_=M*(N+K)*2, flops=2.0*M*N*K)
        db = torch.zeros(N, device=dev)
        timeit(f"wgrad dW[{N},{K}] + bias", lambda: lib.svit_gemm_wgrad_bias(ptr(dY), ptr(X), ptr(dW), ptr(db), M, N, K, N, K, K, SMS, st()), bytes_=M*(N+K)*2, flops=2.0*M*N*K)
if "attn" in which:
    inner = H * 64
    qkv = torch.randn(B, T, 3 * inner, device=dev).bfloat16(); out = torch.empty(B, T, inner, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev); dout = torch.randn(B, T, inner, device=dev).bfloat16()
    dqkv = torch.empty_like(qkv)
    fl = 4.0 * B * H * T * T * 64
    timeit("attn_fwd", lambda: lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, 0.125, st()), flops=fl)
    timeit("attn_bwd", lambda: lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, 0.125, st()), flops=2.5 * fl)
