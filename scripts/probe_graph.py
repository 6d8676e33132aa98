import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import surface_vision_transformers_b200 as svit
dev = torch.device("cuda:0")
cfg = dict(dim=192, depth=12, heads=3, mlp_dim=768, num_patches=320, num_vertices=153)
torch.manual_seed(0)
model = svit.SiT(**cfg).to(dev)
B = 16
x = torch.randn(B, 4, 320, 153, device=dev); y = torch.rand(B, device=dev) * 19 + 26
# ---- inference
model.eval()
with torch.no_grad():
    for _ in range(3): ref = model(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    sx = x.clone()
    with torch.cuda.graph(g):
        out = model(sx)
    g.replay(); torch.cuda.synchronize()
    print("infer graph == eager:", torch.equal(out, ref))
    def timeit(fn, n=50):
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
    print("infer eager ms", timeit(lambda: model(x)), "graph ms", timeit(g.replay))
# ---- training step
model.train()
opt = svit.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=0.0)
def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(model(x).squeeze(), y)
    loss.backward()
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
print("train eager ms", timeit(step, 30))
try:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        sl = step()
    g2.replay(); torch.cuda.synchronize()
    print("train graph ms", timeit(g2.replay, 30), "loss", sl.item())
except Exception as e:
    print("train capture failed:", repr(e)[:300])
