import sys, torch
sys.path.insert(0, "/root/repo")
import surface_vision_transformers_b200 as svit
DEV = torch.device("cuda:0")
cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=20, num_vertices=15)
torch.manual_seed(21)
base = svit.SiT(**cfg).to(DEV)
xs = [torch.randn(8, 4, 20, 15, device=DEV) for _ in range(8)]
ys = [torch.rand(8, device=DEV) * 19 + 26 for _ in range(8)]
def rel(a, b):
    num = sum(((p.detach() - q.detach()).float() ** 2).sum() for p, q in zip(a.parameters(), b.parameters()))
    den = sum((q.detach().float() ** 2).sum() for q in b.parameters())
    return (num / den).sqrt().item()
def go(sync, lr):
    m = svit.SiT(**cfg); m.load_state_dict(base.state_dict()); m.to(DEV)
    o = svit.FusedAdamW(m.parameters(), lr=lr, weight_decay=0.0)
    for k in range(8):
        o.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(m(xs[k]).squeeze(), ys[k]).backward()
        o.step()
        if sync: torch.cuda.synchronize()
    torch.cuda.synchronize()
    return m
for lr in (1e-2, 1e-3):
    a, b, c = go(True, lr), go(True, lr), go(False, lr)
    print("lr", lr, "sync vs sync", rel(a, b), " async vs sync", rel(c, a))
