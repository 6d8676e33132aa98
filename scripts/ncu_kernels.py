"""One launch (after warm-up) of every hot kernel at the SiT-small ico-2 B=256 shapes -- target for
`ncu --set full -k regex:<name> -c 1` captures (profiles/)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surface_vision_transformers_b200 import _lib
from surface_vision_transformers_b200._lib import ptr, vp, check
lib = _lib.load()
dev = torch.device("cuda:0")
B, T, D, H, MLP = 256, 321, 384, 6, 1536
M = B * T
st = vp(torch.cuda.current_stream().cuda_stream)
f = ctypes.c_float
qkv = torch.randn(B, T, 3 * D, device=dev).bfloat16(); out = torch.empty(B, T, D, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, H, T, device=dev); dout = torch.randn(B, T, D, device=dev).bfloat16(); dqkv = torch.empty_like(qkv)
dY = (torch.randn(M, MLP, device=dev) * 0.5).bfloat16(); X = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
dW = torch.zeros(MLP, D, device=dev); db = torch.zeros(MLP, device=dev)
x = torch.randn(M, D, device=dev); g = torch.ones(D, device=dev); bt = torch.zeros(D, device=dev)
a = torch.empty(M, D, device=dev, dtype=torch.bfloat16); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
da = torch.randn(M, D, device=dev).bfloat16(); gi = torch.randn(M, D, device=dev); go = torch.empty_like(gi); g16 = torch.empty_like(a)
dg = torch.zeros(D, device=dev); dbt = torch.zeros(D, device=dev); cs = torch.zeros(D, device=dev)
for _ in range(2):
    check(lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, f(0.125), st), "fwd")
    check(lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, f(0.125), st), "bwd")
    check(lib.svit_gemm_wgrad_bias(ptr(dY), ptr(X), ptr(dW), ptr(db), M, MLP, D, MLP, D, D, 148, st), "wgrad")
    check(lib.svit_layernorm_fwd(ptr(x), ptr(g), ptr(bt), ptr(a), ptr(mean), ptr(rstd), M, D, f(1e-5), st), "ln")
    check(lib.svit_layernorm_bwd(ptr(da), ptr(x), ptr(mean), ptr(rstd), ptr(g), ptr(gi), ptr(go), ptr(g16), ptr(dg), ptr(dbt), ptr(cs), M, D, st), "lnb")
torch.cuda.synchronize()
