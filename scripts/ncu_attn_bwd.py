"""One launch of attn_fwd + attn_bwd at the SiT-small ico-2 shape (target for `ncu -k regex:attn_bwd`)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surface_vision_transformers_b200 import _lib
from surface_vision_transformers_b200._lib import ptr, vp
lib = _lib.load()
B, H, T = int(os.environ.get("B", 256)), 6, 321; inner = 384; dev = torch.device("cuda:0")
qkv = torch.randn(B, T, 3 * inner, device=dev).bfloat16(); out = torch.empty(B, T, inner, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, H, T, device=dev); dout = torch.randn(B, T, inner, device=dev).bfloat16()
dqkv = torch.empty_like(qkv)
st = vp(torch.cuda.current_stream().cuda_stream)
lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, ctypes.c_float(0.125), st)
for _ in range(2):
    lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, ctypes.c_float(0.125), st)
torch.cuda.synchronize()
