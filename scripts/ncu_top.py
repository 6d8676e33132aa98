"""One launch of each of the kernels that dominate the training step, at the benchmark shapes (SiT-small ico-2, per-GPU
batch 256), inside cudaProfilerStart/Stop -- target for

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_top python scripts/ncu_top.py
"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surface_vision_transformers_b200 import _lib
from surface_vision_transformers_b200._lib import ptr, vp, check
lib = _lib.load()
dev = torch.device("cuda:0")
B, T, D, H, MLP = 256, 321, 384, 6, 1536
M = B * T
SMS = torch.cuda.get_device_properties(0).multi_processor_count
st = vp(torch.cuda.current_stream().cuda_stream)
f = ctypes.c_float
bf = torch.bfloat16
qkv = torch.randn(B, T, 3 * D, device=dev).to(bf); out = torch.empty(B, T, D, device=dev, dtype=bf)
lse = torch.zeros(B, H, T, device=dev); dout = torch.randn(B, T, D, device=dev).to(bf); dqkv = torch.empty_like(qkv)
a = (torch.randn(M, D, device=dev) * 0.5).to(bf); w1 = (torch.randn(MLP, D, device=dev) * 0.05).to(bf); b1 = torch.zeros(MLP, device=dev)
h = torch.empty(M, MLP, device=dev, dtype=bf); gp = torch.empty(M, MLP, device=dev, dtype=bf)
w2 = (torch.randn(D, MLP, device=dev) * 0.05).to(bf); b2 = torch.zeros(D, device=dev)
x = torch.randn(M, D, device=dev); xo = torch.empty(M, D, device=dev); an = torch.empty(M, D, device=dev, dtype=bf)
gam = torch.ones(D, device=dev); bet = torch.zeros(D, device=dev); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
g16 = (torch.randn(M, D, device=dev) * 0.5).to(bf); du = torch.empty(M, MLP, device=dev, dtype=bf)
dW = torch.zeros(MLP, D, device=dev); db = torch.zeros(MLP, device=dev)
gi = torch.randn(M, D, device=dev); go = torch.empty_like(gi); g16o = torch.empty(M, D, device=dev, dtype=bf)
dg = torch.zeros(D, device=dev); dbt = torch.zeros(D, device=dev); cs = torch.zeros(D, device=dev)

def once():
    check(lib.svit_attn_fwd(ptr(qkv), ptr(out), ptr(lse), B, H, T, f(0.125), st), "attn_fwd")
    check(lib.svit_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, H, T, f(0.125), st), "attn_bwd")
    # fc1 + bias + GELU and its derivative (EPI_GELU_GRAD = 5)
    check(lib.svit_gemm_tn(ptr(a), ptr(w1), ptr(gp), ptr(h), vp(0), ptr(b1), vp(0), 1, M, MLP, D, D, D, MLP, 5, 0, SMS, st), "fc1")
    # fc2 + bias + residual + next LayerNorm
    check(lib.svit_gemm_ln(ptr(h), ptr(w2), ptr(b2), ptr(x), ptr(xo), ptr(an), ptr(gam), ptr(bet), ptr(mean), ptr(rstd), M, D, MLP,
                           MLP, MLP, f(1e-5), SMS, st), "gemm_ln")
    # dfc2: (g W2) * gelu'  (EPI_MUL = 6), A = g16 [M, D], B = W2^T [MLP, D]
    check(lib.svit_gemm_tn(ptr(g16), ptr(w1), ptr(du), vp(0), ptr(gp), vp(0), vp(0), 1, M, MLP, D, D, D, MLP, 6, 0, SMS, st), "dfc2")
    # dfc1: du W1 (plain store, K = 1536)
    check(lib.svit_gemm_tn(ptr(du), ptr(w2), ptr(an), vp(0), vp(0), vp(0), vp(0), 1, M, D, MLP, MLP, MLP, D, 0, 0, SMS, st), "dfc1")
    check(lib.svit_gemm_wgrad_bias(ptr(du), ptr(a), ptr(dW), ptr(db), M, MLP, D, MLP, D, D, SMS, st), "wgrad")
    check(lib.svit_layernorm_bwd(ptr(an), ptr(x), ptr(mean), ptr(rstd), ptr(gam), ptr(gi), ptr(go), ptr(g16o), ptr(dg), ptr(dbt),
                                 ptr(cs), M, D, st), "ln_bwd")

for _ in range(2):
    once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
