"""GPU probe: full SiT / MPP forward+backward vs the fp32 oracle on identical weights and inputs."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_vision_transformers_b200 as svit
from oracle.sit_oracle import OracleSiT, OracleMPP
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")

def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()

def compare_grads(model, oracle, tol):
    og = dict(oracle.named_parameters()); worst = ("", 0.0); ok = True
    for n, p in model.named_parameters():
        q = og[n]
        if q.grad is None or p.grad is None:
            if (q.grad is None) != (p.grad is None):
                print("  grad presence mismatch", n, q.grad is None, p.grad is None); ok = False
            continue
        e = rel(p.grad, q.grad)
        if e > worst[1]: worst = (n, e)
        if not (e < tol): print(f"  FAIL grad {n}: rel={e:.3e} |ref|={q.grad.norm().item():.3e}"); ok = False
    num = sum(((p.grad - og[n].grad).float() ** 2).sum() for n, p in model.named_parameters() if p.grad is not None)
    den = sum((og[n].grad.float() ** 2).sum() for n, p in model.named_parameters() if p.grad is not None)
    print(f"  grads: global rel={(num / den).sqrt().item():.3e} worst={worst[0]} {worst[1]:.3e}")
    return ok

def run_sit(cfg, B, pool="cls", seed=0):
    torch.manual_seed(seed)
    oracle = OracleSiT(pool=pool, **cfg).to(dev)
    model = svit.SiT(pool=pool, **cfg)
    model.load_state_dict(oracle.state_dict())
    model.to(dev)
    x = torch.randn(B, cfg["num_channels"], cfg["num_patches"], cfg["num_vertices"], device=dev)
    y = torch.rand(B, device=dev) * 19 + 26
    out_o = oracle(x); loss_o = torch.nn.functional.mse_loss(out_o.squeeze(), y); loss_o.backward()
    out_m = model(x); loss_m = torch.nn.functional.mse_loss(out_m.squeeze(), y); loss_m.backward()
    torch.cuda.synchronize()
    enc_o = oracle.encode(x)
    with torch.no_grad():
        xe = oracle.to_patch_embedding(x)
        xe = torch.cat((oracle.cls_token.expand(B, -1, -1), xe), 1) + oracle.pos_embedding
        enc_m = model.transformer(xe)
    e_enc = rel(enc_m, enc_o); e_out = rel(out_m, out_o)
    with torch.no_grad():
        model.eval(); e_inf = rel(model(x), out_o); model.train()
    print(f"SiT {cfg['dim']}/{cfg['depth']} N={cfg['num_patches']} V={cfg['num_vertices']} B={B} pool={pool}: "
          f"enc rel={e_enc:.3e} out rel={e_out:.3e} infer rel={e_inf:.3e} loss {loss_m.item():.5f} vs {loss_o.item():.5f}")
    ok = e_enc < 1e-2 and e_out < 2e-2 and e_inf < 2e-2
    ok &= compare_grads(model, oracle, 2e-2)
    print("  ->", "OK" if ok else "FAIL", flush=True)
    return ok

def run_mpp(cfg, B, seed=0, mask_prob=0.5, replace_prob=0.8, swap_prob=0.02):
    torch.manual_seed(seed)
    K = cfg["num_channels"] * cfg["num_vertices"]
    oracle_sit = OracleSiT(**cfg).to(dev)
    oracle = OracleMPP(oracle_sit, cfg["dim"], K, dev, mask_prob, replace_prob, swap_prob, cfg["num_channels"], cfg["num_vertices"]).to(dev)
    model = svit.SiT(**cfg)
    ssl = svit.masked_patch_pretraining(transformer=model, dim_in=cfg["dim"], dim_out=K, device=dev, mask_prob=mask_prob,
                                        replace_prob=replace_prob, swap_prob=swap_prob, channels=cfg["num_channels"],
                                        num_vertices=cfg["num_vertices"])
    ssl.load_state_dict(oracle.state_dict()); ssl.to(dev)
    x = torch.randn(B, cfg["num_channels"], cfg["num_patches"], cfg["num_vertices"], device=dev)
    from surface_vision_transformers_b200.mpp import draw_masks
    masks = draw_masks(B, cfg["num_patches"], K, dev, mask_prob, replace_prob, swap_prob)
    lo, oo = oracle(x, masks=masks); lo.backward()
    lm, om = ssl(x, masks=masks); lm.backward()
    torch.cuda.synchronize()
    e_out = rel(om, oo)
    print(f"MPP {cfg['dim']}/{cfg['depth']} N={cfg['num_patches']} V={cfg['num_vertices']} B={B}: loss {lm.item():.6f} vs {lo.item():.6f} "
          f"batch_out rel={e_out:.3e}")
    ok = abs(lm.item() - lo.item()) / abs(lo.item()) < 1e-2 and e_out < 2e-2
    ok &= compare_grads(ssl, oracle, 2e-2)
    print("  ->", "OK" if ok else "FAIL", flush=True)
    return ok

if __name__ == "__main__":
    small = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=12, num_vertices=10, num_channels=4, num_classes=1)
    tiny = dict(dim=192, depth=12, heads=3, mlp_dim=768, num_patches=320, num_vertices=153, num_channels=4, num_classes=1)
    sm1 = dict(dim=384, depth=12, heads=6, mlp_dim=1536, num_patches=80, num_vertices=561, num_channels=4, num_classes=1)
    ok = True
    ok &= run_sit(small, 3)
    ok &= run_sit(small, 5, pool="mean")
    ok &= run_mpp(small, 4)
    ok &= run_sit(tiny, 16)
    ok &= run_sit(sm1, 8)
    sm2 = dict(sm1, num_patches=320, num_vertices=153)
    ok &= run_mpp(sm2, 8)
    print("ALL OK" if ok else "SOME FAILED")
