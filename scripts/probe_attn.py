"""GPU probe: fused attention fwd/bwd vs torch fp32 reference on the same bf16 inputs."""
import ctypes, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surface_vision_transformers_b200.build import LIB
lib = ctypes.CDLL(LIB)
lib.svit_last_error.restype = ctypes.c_char_p
vp = ctypes.c_void_p; ci = ctypes.c_int; cf = ctypes.c_float
lib.svit_attn_fwd.argtypes = [vp, vp, vp, ci, ci, ci, cf, vp]
lib.svit_attn_bwd.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, cf, vp]
dev = torch.device("cuda:0")
def ptr(t): return vp(t.data_ptr())
def st(): return vp(torch.cuda.current_stream().cuda_stream)

def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()

def run(B, H, T, bench=False):
    torch.manual_seed(B * 1000 + H * 10 + T)
    inner = H * 64
    qkv = (torch.randn(B, T, 3 * inner, device=dev) * 1.0).bfloat16()
    out = torch.full((B, T, inner), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev)
    scale = 64 ** -0.5
    rc = lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, scale, st())
    assert rc == 0, lib.svit_last_error()
    torch.cuda.synchronize()
    q, k, v = [t.reshape(B, T, H, 64).permute(0, 2, 1, 3).float() for t in qkv.chunk(3, dim=-1)]
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    dots = q @ k.transpose(-1, -2) * scale
    attn = dots.softmax(-1)
    ref = (attn @ v).permute(0, 2, 1, 3).reshape(B, T, inner)
    lse_ref = torch.logsumexp(dots, dim=-1)
    e1, e2 = rel(out, ref), rel(lse, lse_ref)
    ok = e1 < 1e-2 and e2 < 1e-4 and bool(torch.isfinite(out.float()).all())
    print(f"{'OK  ' if ok else 'FAIL'} attn_fwd B={B} H={H} T={T}: out rel={e1:.3e} lse rel={e2:.3e}", flush=True)
    if not ok:
        print(out[0, :3, :8].float().cpu(), ref[0, :3, :8].cpu())
        bad = ((out.float() - ref).abs() > 0.05).nonzero()
        print("#bad", bad.shape[0], bad[:10].tolist(), bad[-5:].tolist())
    # backward
    dout = (torch.randn(B, T, inner, device=dev)).bfloat16()
    delta = torch.zeros(B, H, T, device=dev)
    dqacc = torch.empty(B, T, inner, device=dev)
    dqkv = torch.full((B, T, 3 * inner), float("nan"), device=dev, dtype=torch.bfloat16)
    rc = lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, scale, st())
    assert rc == 0, lib.svit_last_error()
    torch.cuda.synchronize()
    ref.backward(dout.float())
    dref = torch.cat([g.permute(0, 2, 1, 3).reshape(B, T, inner) for g in (q.grad, k.grad, v.grad)], dim=-1)
    names = ["dq", "dk", "dv"]
    okb = True
    for i, n in enumerate(names):
        e = rel(dqkv[..., i * inner:(i + 1) * inner], dref[..., i * inner:(i + 1) * inner])
        o = e < 2e-2 and bool(torch.isfinite(dqkv.float()).all())
        okb &= o
        print(f"{'OK  ' if o else 'FAIL'} attn_bwd {n} B={B} H={H} T={T}: rel={e:.3e}", flush=True)
        if not o:
            g = dqkv[..., i * inner:(i + 1) * inner].float(); r = dref[..., i * inner:(i + 1) * inner]
            print(g[0, :3, :6].cpu(), r[0, :3, :6].cpu())
            bad = ((g - r).abs() > 0.1 * r.abs().max()).nonzero()
            print("#bad", bad.shape[0], bad[:10].tolist(), bad[-5:].tolist())
    if bench:
        for name, f in (("fwd", lambda: lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, scale, st())),
                        ("bwd", lambda: lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, scale, st()))):
            for _ in range(3): f()
            a = torch.cuda.Event(enable_timing=True); bb = torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): f()
            bb.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(bb) / 10
            fl = 4 * B * H * T * T * 64 * (1 if name == "fwd" else 2.5)
            print(f"bench attn_{name} B={B} H={H} T={T}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    return ok and okb

if __name__ == "__main__":
    allok = True
    for (B, H, T) in [(1, 1, 128), (2, 3, 81), (2, 2, 321), (3, 6, 321), (1, 1, 21), (2, 2, 384), (1, 2, 200)]:
        allok &= run(B, H, T)
    print("ALL OK" if allok else "SOME FAILED")
    run(256, 6, 321, bench=True)
