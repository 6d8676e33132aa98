"""Multi-GPU correctness of the data-parallel path, run under torchrun with N ranks on one node:
every rank steps on its shard of a global batch through DataParallel + FusedAdamW; rank 0 also trains a single-process
replica on the WHOLE batch.  With a mean loss the all-reduced average of the shard gradients equals the full-batch
gradient, so after a few steps the replicas must hold the same parameters as the single-process model (bf16 noise)
and be bit-identical to each other.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/check_ddp.py
"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_vision_transformers_b200 as svit

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl")
cfg = dict(dim=192, depth=4, heads=3, mlp_dim=768, num_patches=80, num_vertices=45)
per_rank, steps = 8, 4
torch.manual_seed(0)                                   # same initial weights and the same global data on every rank
ref = svit.SiT(**cfg).to(dev)
model = svit.SiT(**cfg)
model.load_state_dict(ref.state_dict())
model.to(dev)
xs = [torch.randn(per_rank * world, 4, 80, 45, device=dev) for _ in range(steps)]
ys = [torch.rand(per_rank * world, device=dev) * 19 + 26 for _ in range(steps)]
ddp = svit.DataParallel(model)
opt = svit.FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.0)
opt_ref = svit.FusedAdamW(ref.parameters(), lr=1e-3, weight_decay=0.0)
sl = slice(rank * per_rank, (rank + 1) * per_rank)
for x, y in zip(xs, ys):
    opt.zero_grad(set_to_none=True)
    torch.nn.functional.mse_loss(ddp(x[sl]).squeeze(-1), y[sl]).backward()
    opt.step()
    opt_ref.zero_grad(set_to_none=True)
    torch.nn.functional.mse_loss(ref(x).squeeze(-1), y).backward()
    opt_ref.step()
torch.cuda.synchronize()
flat = model._flat.detach().clone()
num = (flat - ref._flat.detach()).double().pow(2).sum().sqrt().item()
den = ref._flat.detach().double().pow(2).sum().sqrt().item()
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(g, gathered[0]) for g in gathered)
if rank == 0:
    print(f"ddp x{world}: replicas bit-identical = {same}; parameters vs single-process full-batch training: rel-L2 {num / den:.3e}")
    assert same and num / den < 1e-4
dist.destroy_process_group()
