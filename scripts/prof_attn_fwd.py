"""clock64 timeline of one CTA of attn_fwd_kernel (SVIT_ATTN_DEBUG=8) at the SiT-small ico-2 shape."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SVIT_ATTN_DEBUG"] = "8"
from surface_vision_transformers_b200 import _lib
from surface_vision_transformers_b200._lib import ptr, vp
lib = ctypes.CDLL(_lib._build.LIB)
B, H, T = 256, 6, int(os.environ.get("T", 321)); inner = 384; dev = torch.device("cuda:0")
qkv = torch.randn(B, T, 3 * inner, device=dev).bfloat16(); out = torch.empty(B, T, inner, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, H, T, device=dev)
st = vp(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, ctypes.c_float(0.125), st)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 256)()
lib.svit_debug_attn_prof(buf, 256)
v = list(buf); t0 = v[0]
print("control thread: K/V landed", v[1] - t0)
for i in range(3):
    r = [v[10 + i * 10 + k] - t0 if v[10 + i * 10 + k] else None for k in range(8)]
    print(f" q-block {i}: Q landed {r[0]} | S retired {r[1]} | ctrl saw P {r[2]} | PV issued {r[3]} || softmax warp 0: exps done {r[4]} | "
          f"sums exchanged {r[5]} | P in TMEM {r[6]} | O ready {r[7]}")
