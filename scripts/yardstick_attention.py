"""Yardstick (NOT part of the product path): library attention kernels on the SiT-small ico-2 shape, to place the
hand-written tcgen05 kernels.  flash_attn 2.x (mma.sync, sm_80-style) and torch SDPA backends."""
import torch, time
B, H, T, D = 256, 6, 321, 64
dev = torch.device("cuda:0")
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
fl = 4.0 * B * H * T * T * D
try:
    from flash_attn import flash_attn_func
    q, k, v = [torch.randn(B, T, H, D, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(3)]
    us = timeit(lambda: flash_attn_func(q, k, v))
    o = flash_attn_func(q, k, v); do = torch.randn_like(o)
    usb = timeit(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True))
    print(f"flash_attn {__import__('flash_attn').__version__}: fwd {us:.0f} us ({fl/us/1e6:.0f} TF/s)  bwd {usb:.0f} us ({2.5*fl/usb/1e6:.0f} TF/s)")
except Exception as e:
    print("flash_attn unavailable:", repr(e)[:200])
from torch.nn.attention import sdpa_kernel, SDPBackend
q, k, v = [torch.randn(B, H, T, D, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(3)]
for be in (SDPBackend.FLASH_ATTENTION, SDPBackend.CUDNN_ATTENTION, SDPBackend.EFFICIENT_ATTENTION):
    try:
        with sdpa_kernel(be):
            f = lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)
            us = timeit(f)
            o = f(); do = torch.randn_like(o)
            usb = timeit(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True))
        print(f"torch SDPA {be.name}: fwd {us:.0f} us ({fl/us/1e6:.0f} TF/s)  bwd {usb:.0f} us ({2.5*fl/usb/1e6:.0f} TF/s)")
    except Exception as e:
        print(f"torch SDPA {be.name}: unavailable ({repr(e)[:120]})")
