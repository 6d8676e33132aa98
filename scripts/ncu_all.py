"""One launch of EVERY kernel of the training step at the benchmark shapes (SiT-small ico-2, per-GPU batch 256), for

    ncu --section LaunchStats --section Occupancy --section SpeedOfLight --section MemoryWorkloadAnalysis \
        --section ComputeWorkloadAnalysis --section WarpStateStats \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sass__inst_executed_local_loads,sass__inst_executed_local_stores \
        --clock-control none --profile-from-start off \
        -o gpurun_out/r02_all python scripts/ncu_all.py

(the sections that hold duration, DRAM bytes, pipe utilisation, occupancy and stall reasons: ~10 replay passes per launch
instead of the ~40 of --set full, which took 20 GPU-minutes for this script; the attention and GEMM kernels additionally
have --set full captures with source, scripts/bench_attn.py once / scripts/ncu_kernels.py)

A 2-block SiT-small (same kernels and shapes as the 12-block model, 1/6 of the launches) runs one un-profiled warm-up
iteration, then -- inside cudaProfilerStart/Stop -- one iteration of: weight-shadow refresh, forward, fused criterion,
backward, fused AdamW; the same through the MPP module; one SGD step; the raw-mesh gather.  scripts/ncu_summary.py
condenses the report into profiles/."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_vision_transformers_b200 as svit
dev = torch.device("cuda:0")
B = int(os.environ.get("B", 256))
cfg = dict(dim=384, depth=2, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153, num_channels=4, num_classes=1)
torch.manual_seed(0)
model = svit.SiT(**cfg).to(dev)
opt = svit.FusedAdamW(model.parameters(), lr=1e-5, weight_decay=0.0)
ssl = svit.masked_patch_pretraining(transformer=svit.SiT(**dict(cfg, depth=1)), dim_in=384, dim_out=612, device=dev, mask_prob=0.5,
                                    replace_prob=0.8, swap_prob=0.02, channels=4, num_vertices=153).to(dev)
opt2 = svit.FusedAdamW(ssl.parameters(), lr=1e-4, weight_decay=0.0)
sgd_model = svit.SiT(**dict(cfg, depth=1)).to(dev)
sgd = svit.FusedSGD(sgd_model.parameters(), lr=1e-4, momentum=0.9)
x = torch.randn(B, 4, 320, 153, device=dev)
y = torch.rand(B, device=dev) * 19 + 26
mesh = torch.randn(64, 4, 40962, device=dev)
table = svit.load_index_table(2, dev)

def iteration():
    opt.zero_grad(set_to_none=True)
    loss = svit.regression_loss(model(x), y)
    loss.backward()
    opt.step()
    opt2.zero_grad(set_to_none=True)
    l2, _ = ssl(x)
    l2.backward()
    opt2.step()
    sgd.zero_grad(set_to_none=True)
    svit.regression_loss(sgd_model(x[:32]), y[:32]).backward()
    sgd.step()
    svit.gather_patches(mesh, table)

iteration()
torch.cuda.synchronize()
torch.cuda.profiler.start()
iteration()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
