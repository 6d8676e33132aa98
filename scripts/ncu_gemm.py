"""A few launches of one TN GEMM shape (target for `ncu -k regex:gemm_tn`): fc1 + bias + GELU at SiT-small ico-2 B=256."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surface_vision_transformers_b200 import _lib
from surface_vision_transformers_b200._lib import ptr, vp, check
lib = _lib.load()
dev = torch.device("cuda:0")
M, N, K, mode = 256 * 321, int(os.environ.get("N", 1536)), int(os.environ.get("K", 384)), int(os.environ.get("MODE", 1))
A = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
bias = torch.zeros(N, device=dev); o = torch.empty(M, N, device=dev, dtype=torch.bfloat16); o2 = torch.empty_like(o)
aux = torch.randn(M, N, device=dev).bfloat16()
st = vp(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    check(lib.svit_gemm_tn(ptr(A), ptr(W), ptr(o), ptr(o2), ptr(aux), ptr(bias), vp(0), 1, M, N, K, K, K, N, mode, 0, 148, st), "gemm")
torch.cuda.synchronize()
