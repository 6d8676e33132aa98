"""Cost of dropout > 0 (separate HBM-bound passes, csrc/dropout.cu) on the SiT-small ico-2 training step, B = 256."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_vision_transformers_b200 as svit
dev = torch.device("cuda:0")
cfg = dict(dim=384, depth=12, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153)
B = int(os.environ.get("B", 256))
x = torch.randn(B, 4, 320, 153, device=dev); y = torch.rand(B, device=dev) * 19 + 26
for p in (0.0, 0.1):
    torch.manual_seed(0)
    model = svit.SiT(**cfg, dropout=p, emb_dropout=p).to(dev)
    opt = svit.FusedAdamW(model.parameters(), lr=1e-4)
    def step():
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(model(x).squeeze(), y).backward()
        opt.step()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"dropout={p}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} samples/s")
