"""Attention micro-benchmark at the SiT-small ico-2 shape (B=256, H=6, T=321, d=64), CUDA-event timed:
forward and backward on N(0,1) logits and on PEAKY logits (qkv x 8: rows whose shift-by-key-0 softmax overflows take
the max-shift "safe mode" redo of the forward kernel).  `python scripts/bench_attn.py once` = warm-up + one launch
of each kernel (target for ncu)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surface_vision_transformers_b200 import _lib
from surface_vision_transformers_b200._lib import ptr, vp, check
lib = _lib.load()
dev = torch.device("cuda:0")
B, H, T = int(os.environ.get("B", 256)), 6, int(os.environ.get("T", 321)); inner = H * 64
st = lambda: vp(torch.cuda.current_stream().cuda_stream)
f = ctypes.c_float(0.125)
once = "once" in sys.argv[1:]
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
fl = 4.0 * B * H * T * T * 64
for name, mul in (("N(0,1) logits", 1.0), ("peaky logits (qkv x 8)", 8.0)):
    torch.manual_seed(0)
    qkv = (torch.randn(B, T, 3 * inner, device=dev) * mul).bfloat16()
    out = torch.empty(B, T, inner, device=dev, dtype=torch.bfloat16); lse = torch.zeros(B, H, T, device=dev)
    dout = torch.randn(B, T, inner, device=dev).bfloat16(); dqkv = torch.empty_like(qkv)
    fwd = lambda: check(lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, f, st()), "fwd")
    bwd = lambda: check(lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, f, st()), "bwd")
    if once:
        for _ in range(2): fwd(); bwd()
        torch.cuda.synchronize()
        break
    uf, ub = timeit(fwd), timeit(bwd)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    print(f"{name:26s} attn_fwd {uf:7.1f} us ({fl/uf/1e6:5.0f} TF/s)   attn_bwd {ub:7.1f} us ({2.5*fl/ub/1e6:5.0f} TF/s)", flush=True)
