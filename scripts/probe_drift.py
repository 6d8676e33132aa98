"""Where does the parameter drift between two runs of the same 8 AdamW steps come from? (debug probe)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_vision_transformers_b200 as svit
DEV = torch.device("cuda:0")
cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=20, num_vertices=15)
torch.manual_seed(21)
base = svit.SiT(**cfg).to(DEV)
xs = [torch.randn(8, 4, 20, 15, device=DEV) for _ in range(8)]
ys = [torch.rand(8, device=DEV) * 19 + 26 for _ in range(8)]
def run(sync, steps=8, poison=False):
    m = svit.SiT(**cfg); m.load_state_dict(base.state_dict()); m.to(DEV)
    o = svit.FusedAdamW(m.parameters(), lr=1e-3, weight_decay=0.0)
    snaps = []
    for k in range(steps):
        if poison:   # fill the allocator's free blocks with NaN bit patterns
            junk = [torch.full((n,), float("nan"), device=DEV) for n in (1 << 22, 1 << 20, 1 << 18, 1 << 16, 1 << 14)]
            del junk
        o.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(m(xs[k]).squeeze(), ys[k]).backward()
        snaps.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone())
        o.step()
        if sync: torch.cuda.synchronize()
    torch.cuda.synchronize()
    return m._flat.clone(), snaps
def rel(a, b): return ((a - b).norm() / b.norm()).item()
ref, gref = run(True)
for name, kw in (("sync again", dict(sync=True)), ("no sync", dict(sync=False)), ("sync + poisoned free blocks", dict(sync=True, poison=True)),
                 ("no sync again", dict(sync=False))):
    p, g = run(**kw)
    print(f"{name:30s} params rel {rel(p, ref):.2e}   grads per step " + " ".join(f"{rel(a, b):.1e}" for a, b in zip(g, gref)), flush=True)
