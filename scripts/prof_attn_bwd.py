import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SVIT_ATTN_DEBUG"] = os.environ.get("SVIT_ATTN_DEBUG", "4")
from surface_vision_transformers_b200 import _lib
from surface_vision_transformers_b200._lib import ptr, vp
lib = ctypes.CDLL(_lib._build.LIB)
B, H, T = 256, 6, 321; inner = 384; dev = torch.device("cuda:0")
qkv = torch.randn(B, T, 3 * inner, device=dev).bfloat16(); out = torch.empty(B, T, inner, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, H, T, device=dev); dout = torch.randn(B, T, inner, device=dev).bfloat16()
delta = torch.zeros(B, H, T, device=dev); dqkv = torch.empty_like(qkv); dqacc = torch.empty(B, T, inner, device=dev)
st = vp(torch.cuda.current_stream().cuda_stream)
lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, ctypes.c_float(0.125), st)
for _ in range(3):
    lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(delta), ptr(dqacc), ptr(dqkv), B, H, T, ctypes.c_float(0.125), st)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 256)()
lib.svit_debug_attn_prof(buf, 256)
v = list(buf); t0 = v[0]
def rel(i): return v[i] - t0 if v[i] else None
print("MMA thread: start 0, kv+q landed", rel(1), "a0,b0 issued", rel(2))
for i in range(3):
    print(f" pair {i}: wait p_full {rel(10+i*10)}->{rel(11+i*10)}  c issued {rel(12+i*10)}  wait ds_full {rel(13+i*10)}->{rel(14+i*10)}  d issued {rel(15+i*10)}")
print("compute thread 0:")
for i in range(3):
    print(f" pair {i}: wait s_full {rel(100+i*10)}->{rel(101+i*10)}  P written {rel(102+i*10)}  wait dp_full {rel(103+i*10)}->{rel(104+i*10)}  dS written {rel(105+i*10)}")
print(" dq readout done", rel(150), " store done", rel(151))
