"""fit() / fit_mpp() under torchrun on real GPUs: every rank trains on its slice of each shuffled batch, the epoch
statistics are all-reduced, rank 0 writes the checkpoints.  Prints the (rank-identical) history.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 scripts/check_fit_ddp.py
"""
import os, sys, tempfile, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_vision_transformers_b200 as svit

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl")
root = os.path.join(tempfile.gettempdir(), "svit_check_fit")
if rank == 0:
    os.makedirs(root, exist_ok=True)
    rng = np.random.default_rng(0)
    for split, n in (("train", 64), ("validation", 16)):
        x = rng.standard_normal((n, 4, 80, 45))
        np.save(os.path.join(root, f"{split}_data.npy"), x)
        np.save(os.path.join(root, f"{split}_labels.npy"), 35.0 + 3.0 * x[:, 0].mean(axis=(1, 2)) + 0.1 * rng.standard_normal(n))
dist.barrier()
train, val = svit.PatchedNpyDataset(root, "train"), svit.PatchedNpyDataset(root, "validation")
cfg = dict(dim=192, depth=3, heads=3, mlp_dim=768, num_patches=80, num_vertices=45)
torch.manual_seed(0)
model = svit.SiT(**cfg).to(dev)
res = svit.fit(model, svit.FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.0), train, val, epochs=3, batch_size=16,
               val_epoch=1, device=dev, save_dir=os.path.join(root, "out"), save_ckpt=True, seed=1)
h = res["history"]
t = torch.tensor(h["train_loss"] + [v for _, v in h["val_mae"]], device=dev, dtype=torch.float64)
g = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(g, t)
kw = dict(mask_prob=0.5, replace_prob=0.8, swap_prob=0.02, channels=4, num_vertices=45)
ssl = svit.masked_patch_pretraining(transformer=svit.SiT(**cfg), dim_in=192, dim_out=180, device=dev, **kw).to(dev)
res2 = svit.fit_mpp(ssl, svit.FusedAdamW(ssl.parameters(), lr=1e-3, weight_decay=0.0), train, val, epochs=2, batch_size=16,
                    val_epoch=1, device=dev, save_dir=os.path.join(root, "out_mpp"), seed=2)
if rank == 0:
    print("fit x%d: train loss %s val MAE %s best epoch %s; history identical on all ranks: %s" % (
        world, [round(v, 3) for v in h["train_loss"]], [round(v, 3) for _, v in h["val_mae"]], res["best_epoch"],
        all(torch.equal(a, g[0]) for a in g)))
    print("fit_mpp x%d: train loss %s val loss %s; checkpoints: %s" % (
        world, [round(v, 4) for v in res2["history"]["train_loss"]], [round(v, 4) for _, v in res2["history"]["val_loss"]],
        sorted(os.listdir(os.path.join(root, "out_mpp")))))
    assert all(np.isfinite(h["train_loss"])) and h["train_loss"][-1] < h["train_loss"][0]
    assert os.path.exists(os.path.join(root, "out", "checkpoint.pth"))
dist.destroy_process_group()
