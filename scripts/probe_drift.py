"""Why do two runs of the same 8 AdamW steps sometimes end 3e-5 apart?  (debug probe, GPU)

Hypothesis: the fp32 atomics of the weight-gradient kernels make gradients differ by ~4e-8 from run to run; AdamW turns
that into parameter differences of ~1e-9; whenever one of those straddles a bf16 rounding boundary a weight SHADOW flips by
one bf16 ulp (4e-3 of that weight) and the next step's gradients differ by ~5e-5 -- a chaotic amplification, not a race.
The probe checks the three links separately:
  (1) forward + backward at a FIXED state, repeated: gradients may only differ at the atomics level;
  (2) trials of step 1 + a second forward/backward: the second gradients differ by > 1e-6 exactly in the trials in which at
      least one bf16-rounded parameter differs from trial 0;
  (3) the same 8 steps in fp32 check mode (no bf16 rounding anywhere): every trial ends within 1e-7.
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_vision_transformers_b200 as svit
DEV = torch.device("cuda:0")
cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=20, num_vertices=15)
torch.manual_seed(21)
base = svit.SiT(**cfg).to(DEV)
xs = [torch.randn(8, 4, 20, 15, device=DEV) for _ in range(8)]
ys = [torch.rand(8, device=DEV) * 19 + 26 for _ in range(8)]


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def fresh(check=False):
    m = svit.SiT(**cfg); m.load_state_dict(base.state_dict()); m.to(DEV)
    if check:
        m.set_check_mode(True)
    return m, svit.FusedAdamW(m.parameters(), lr=1e-3, weight_decay=0.0)


def grads(m):
    return torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone()


def fb(m, o, k):
    o.zero_grad(set_to_none=True)
    torch.nn.functional.mse_loss(m(xs[k]).squeeze(), ys[k]).backward()
    return grads(m)


print("(1) forward + backward at a fixed state, 12 repeats")
m, o = fresh()
g0 = fb(m, o, 0)
d = [rel(fb(m, o, 0), g0) for _ in range(12)]
print("    gradient rel-L2 vs repeat 0: max %.2e  (bitwise equal in %d of 12)" % (max(d), sum(x == 0.0 for x in d)))

print("(2) step 1, then a second forward + backward; 10 trials against trial 0")
ref = None
for t in range(10):
    m, o = fresh()
    g1 = fb(m, o, 0)
    o.step()
    p1 = m._flat.clone()
    g2 = fb(m, o, 1)
    torch.cuda.synchronize()
    if ref is None:
        ref = (g1, p1, g2)
        continue
    flips = int((p1.bfloat16() != ref[1].bfloat16()).sum())
    print("    trial %d: grads(1) %.1e   params after step 1 %.1e   bf16-rounded parameters that differ %4d   grads(2) %.1e"
          % (t, rel(g1, ref[0]), rel(p1, ref[1]), flips, rel(g2, ref[2])))

for check in (False, True):
    print("(3) 8 AdamW steps, %s, 6 trials against trial 0" % ("fp32 check mode" if check else "tensor-core path"))
    ref = None
    for t in range(6):
        m, o = fresh(check)
        for k in range(8):
            fb(m, o, k)
            o.step()
        torch.cuda.synchronize()
        p = m._flat.clone()
        if ref is None:
            ref = p
        else:
            print("    trial %d: parameters rel-L2 %.2e" % (t, rel(p, ref)))
