"""GPU probe: tcgen05 GEMM kernels vs torch fp32 matmul of the same bf16 operands. Prints diagnostics."""
import ctypes, sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surface_vision_transformers_b200.build import LIB

lib = ctypes.CDLL(LIB)
lib.svit_last_error.restype = ctypes.c_char_p
vp = ctypes.c_void_p
lib.svit_gemm_tn.argtypes = [vp] * 7 + [ctypes.c_int] * 10 + [vp]
lib.svit_gemm_wgrad.argtypes = [vp, vp, vp] + [ctypes.c_int] * 7 + [vp]
dev = torch.device("cuda:0")
SMS = torch.cuda.get_device_properties(0).multi_processor_count
print("device", torch.cuda.get_device_name(0), "SMs", SMS, flush=True)


def ptr(t):
    return vp(t.data_ptr()) if t is not None else vp(0)


def gelu(x):
    return torch.nn.functional.gelu(x)


def dgelu(x):
    cdf = 0.5 * (1 + torch.erf(x * 0.7071067811865476))
    pdf = 0.3989422804014327 * torch.exp(-0.5 * x * x)
    return cdf + x * pdf


def report(name, got, ref):
    got = got.float(); ref = ref.float()
    err = (got - ref).norm() / (ref.norm() + 1e-30)
    mx = (got - ref).abs().max()
    ok = bool(err < 1e-2) and bool(torch.isfinite(got).all())
    print(f"{'OK  ' if ok else 'FAIL'} {name}: rel_l2={err.item():.3e} max_abs={mx.item():.3e}", flush=True)
    if not ok:
        print(" got[:4,:8]\n", got[:4, :8].cpu())
        print(" ref[:4,:8]\n", ref[:4, :8].cpu())
        bad = ((got - ref).abs() > 1e-2 * ref.abs().max()).nonzero()
        print(" #bad", bad.shape[0], "first bad idx", bad[:8].tolist())
    return ok


def run_tn(M, N, K, mode, out_f32, bias=True, rowtab=False):
    torch.manual_seed(M + N + K + mode)
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    odt = torch.float32 if out_f32 else torch.bfloat16
    out = torch.full((M, N), float("nan"), device=dev, dtype=odt)
    out2 = torch.full((M, N), float("nan"), device=dev, dtype=odt) if mode == 1 else None
    aux = (torch.randn(M, N, device=dev)).to(odt) if mode in (2, 3) else None
    period = 7
    rt = torch.randn(period, N, device=dev) if rowtab else None
    rc = lib.svit_gemm_tn(ptr(A), ptr(B), ptr(out), ptr(out2), ptr(aux), ptr(b), ptr(rt), period, M, N, K, K, K, N,
                          mode, int(out_f32), SMS, vp(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        print("FAIL launch rc", rc, lib.svit_last_error().decode()); return False
    torch.cuda.synchronize()
    acc = A.float() @ B.float().t()
    if bias: acc = acc + b
    name = f"tn M={M} N={N} K={K} mode={mode} f32={out_f32} bias={bias} rowtab={rowtab}"
    if mode == 0:
        if rowtab: acc = acc + rt[torch.arange(M, device=dev) % period]
        return report(name, out, acc)
    if mode == 1:
        ok1 = report(name + " [pre]", out, acc)
        ok2 = report(name + " [gelu]", out2, gelu(acc.bfloat16().float()))
        return ok1 and ok2
    if mode == 2:
        return report(name, out, acc + aux.float())
    if mode == 3:
        return report(name, out, acc * dgelu(aux.float()))


def run_wgrad(M, N, K):
    torch.manual_seed(M + N + K)
    dY = (torch.randn(M, N, device=dev) * 0.5).bfloat16()
    X = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    dW = torch.zeros(N, K, device=dev)
    rc = lib.svit_gemm_wgrad(ptr(dY), ptr(X), ptr(dW), M, N, K, N, K, K, SMS, vp(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        print("FAIL launch rc", rc, lib.svit_last_error().decode()); return False
    torch.cuda.synchronize()
    ref = dY.float().t() @ X.float()
    return report(f"wgrad M={M} N={N} K={K}", dW, ref)


def bench_tn(M, N, K, mode, out_f32, iters=20):
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    b = torch.randn(N, device=dev)
    odt = torch.float32 if out_f32 else torch.bfloat16
    out = torch.empty((M, N), device=dev, dtype=odt)
    out2 = torch.empty((M, N), device=dev, dtype=odt) if mode == 1 else None
    aux = torch.randn(M, N, device=dev).to(odt) if mode in (2, 3) else None
    st = vp(torch.cuda.current_stream().cuda_stream)
    f = lambda: lib.svit_gemm_tn(ptr(A), ptr(B), ptr(out), ptr(out2), ptr(aux), ptr(b), vp(0), 1, M, N, K, K, K, N, mode, int(out_f32), SMS, st)
    for _ in range(3): f()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"bench tn M={M} N={N} K={K} mode={mode} f32={out_f32}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters): torch.matmul(A, B.t())
    t1.record(); torch.cuda.synchronize()
    ms2 = t0.elapsed_time(t1) / iters
    print(f"      cublas bf16 same shape: {ms2*1e3:.1f} us  {2*M*N*K/ms2/1e9:.1f} TFLOP/s", flush=True)


def bench_wgrad(M, N, K, iters=20):
    dY = (torch.randn(M, N, device=dev) * 0.5).bfloat16()
    X = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    dW = torch.zeros(N, K, device=dev)
    st = vp(torch.cuda.current_stream().cuda_stream)
    f = lambda: lib.svit_gemm_wgrad(ptr(dY), ptr(X), ptr(dW), M, N, K, N, K, K, SMS, st)
    for _ in range(3): f()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"bench wgrad M={M} N={N} K={K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    allok = True
    # smallest: single tile, single k-block
    allok &= run_tn(128, 192, 64, 0, True, bias=False)
    allok &= run_tn(128, 192, 64, 0, False, bias=False)
    allok &= run_tn(128, 192, 384, 0, True)
    allok &= run_tn(321 * 4, 1152, 384, 0, False, bias=False)
    allok &= run_tn(321 * 4, 384, 384, 2, True)
    allok &= run_tn(321 * 4, 1536, 384, 1, False)
    allok &= run_tn(321 * 4, 384, 1536, 2, True)
    allok &= run_tn(321 * 4, 1536, 384, 3, False, bias=False)
    allok &= run_tn(321 * 4, 384, 640, 0, True, rowtab=True)
    allok &= run_tn(321 * 16, 612, 384, 0, True)
    allok &= run_tn(321 * 64, 1152, 384, 0, False, bias=False)  # multi-wave persistent
    allok &= run_tn(100, 100, 72, 0, True)  # ragged everything
    allok &= run_wgrad(64, 128, 192)
    allok &= run_wgrad(128, 128, 192)
    allok &= run_wgrad(321 * 4, 1152, 384)
    allok &= run_wgrad(321 * 64, 384, 1536)
    allok &= run_wgrad(1000, 104, 72)
    print("ALL OK" if allok else "SOME FAILED", flush=True)
    if allok or "--bench" in sys.argv:
        M = 321 * 256
        bench_tn(M, 1152, 384, 0, False)
        bench_tn(M, 384, 384, 2, True)
        bench_tn(M, 1536, 384, 1, False)
        bench_tn(M, 384, 1536, 2, True)
        bench_tn(M, 1536, 384, 3, False)
        bench_wgrad(M, 1152, 384)
        bench_wgrad(M, 1536, 384)
        bench_wgrad(M, 384, 1536)
