"""clock64 timeline of one CTA of attn_bwd_kernel (SVIT_ATTN_DEBUG=4) at the SiT-small ico-2 shape."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SVIT_ATTN_DEBUG"] = os.environ.get("SVIT_ATTN_DEBUG", "4")
from surface_vision_transformers_b200 import _lib
from surface_vision_transformers_b200._lib import ptr, vp
lib = ctypes.CDLL(_lib._build.LIB)
B, H, T = 256, 6, int(os.environ.get("T", 321)); inner = 384; dev = torch.device("cuda:0")
qkv = torch.randn(B, T, 3 * inner, device=dev).bfloat16(); out = torch.empty(B, T, inner, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, H, T, device=dev); dout = torch.randn(B, T, inner, device=dev).bfloat16()
dqkv = torch.empty_like(qkv)
st = vp(torch.cuda.current_stream().cuda_stream)
lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, ctypes.c_float(0.125), st)
for _ in range(3):
    lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, ctypes.c_float(0.125), st)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 256)()
lib.svit_debug_attn_prof(buf, 256)
v = list(buf); t0 = v[0]
def rel(i): return v[i] - t0 if v[i] else None
print("MMA thread X: start 0, first operands landed", rel(1), " last MMA issued", rel(2), " dQ stored", rel(3))
print("compute warp: start", rel(90), "delta loaded", rel(91), "after barrier", rel(92), "S(0) ready", rel(93), "P(0) stored", rel(94))
for p in range(8):
    print(f" step {p}: X wait p_full {rel(10+p*4)}->{rel(11+p*4)}  Y wait ds_full {rel(12+p*4)}->{rel(13+p*4)} |"
          f" compute: top {rel(100+p*4)} waits done {rel(101+p*4)} dS stored {rel(102+p*4)} P stored {rel(103+p*4)}")
for j in range(1, 4):
    print(f" copy-out of key block {j-1}: {rel(80+2*j)} -> {rel(81+2*j)}")
print(" inside copy-out 0: accumulators final", rel(70), "staged", rel(71), "after barrier", rel(72))
print("tail: last copy-out", rel(95), "->", rel(96), " dq_full", rel(97))
for p in (1, 2):
    print(f"step {p}: dS stored per compute warp (warps 4..15):", [rel(200 + (p - 1) * 16 + w) for w in range(4, 16)])
