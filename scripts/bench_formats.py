"""Measurements for the SURVEY 8(f) rows: on-device preprocessing (z-score + patch gather, HBM-bound) and one epoch of
fit() from pinned host arrays (SiT-small ico-2), next to bench.py's end-to-end number."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surface_vision_transformers_b200 as svit
dev = torch.device("cuda:0")

# ---- preprocessing: (2S, 4, 40962) raw hemispheres -> (2S, 4, 320, 153) patches
S2 = 128
table = svit.load_index_table(2, dev)
hemis = torch.randn(S2, 4, 40962, device=dev)
means, stds = [1.15, 0.037, 1.0, 0.07], [0.41, 0.19, 0.39, 4.05]
for _ in range(2):
    out = svit.preprocess_meshes(hemis, means, stds, table)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    out = svit.preprocess_meshes(hemis, means, stds, table)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
norm = ((hemis.double() - 1) / 2).float()
a.record()
for _ in range(5):
    g = svit.gather_patches(norm, table)
b.record(); torch.cuda.synchronize()
gms = a.elapsed_time(b) / 5
gbytes = norm.numel() * 4 + g.numel() * 4
print(f"preprocess_meshes (float64 z-score + gather + L/R order) {S2} hemispheres: {ms:.2f} ms = {S2 / ms * 1e3:.0f} hemispheres/s")
print(f"  gather kernel alone: {gms * 1e3:.0f} us, {gbytes / gms / 1e6:.0f} GB/s (read mesh + write patches)")

# ---- one epoch of fit() from pinned host arrays
class Synth:
    def __init__(self, n):
        self.data = torch.randn(n, 4, 320, 153).pin_memory()
        self.labels = (torch.rand(n) * 19 + 26).pin_memory()
    def __len__(self): return self.data.shape[0]
    batches = svit.PatchedNpyDataset.batches
n = 2048
ds = Synth(n)
model = svit.SiT(dim=384, depth=12, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153).to(dev)
opt = svit.FusedAdamW(model.parameters(), lr=1e-5, weight_decay=0.0)
svit.fit(model, opt, ds, epochs=1, batch_size=256, device=dev)          # warm-up epoch
torch.cuda.synchronize(); t0 = time.perf_counter()
res = svit.fit(model, opt, ds, epochs=2, batch_size=256, device=dev)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"fit(): 2 epochs x {n} samples (shuffled, pinned host -> DevicePrefetcher -> step, device-side statistics): "
      f"{2 * n / dt:.0f} samples/s  (train loss {res['history']['train_loss'][-1]:.3f})")
