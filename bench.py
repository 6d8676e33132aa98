#!/usr/bin/env python
"""bench.py -- SiT train samples/s on B200 (BASELINE.json metric) for the B200-native hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--workload NAME]

N > 1 is launched by the driver as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
(one rank per GPU, NCCL).  A "step" is one full training iteration of the workload: forward, MSE loss, backward,
gradient all-reduce (N > 1) and a fused AdamW update, on one synthetic batch per GPU (weak scaling).

One JSON line is printed by rank 0:
  value        whole-job samples/s with the inputs already resident in HBM (CUDA-event timed, max over ranks)
  e2e          same metric through the public API with HOST (pinned) inputs: H2D copy of every step's batch
               (svit.DevicePrefetcher: copy of batch i+1 under the kernels of batch i) and a D2H read of every step's
               loss inside the timed region
  roofline     the kernel with the largest share of the step (fused attention backward when training, forward when
               inferring) timed live with CUDA events, whole-step tensor fraction next to it; roofline_gemm = the same
               for the largest GEMM launch (MLP up-projection + bias + GELU epilogue)
  cpu_baseline the oracle port of the reference (fp32 PyTorch on the host cores) on a bounded sample
--impl reference times that CPU path as its own arm (rank 0 only).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the headline metric is quoted on
    "sit_small_ico2_scan_age_train": dict(
        model=dict(dim=384, depth=12, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153, num_channels=4,
                   num_classes=1, dim_head=64), kind="train", gflop_per_sample=46.895),
    "sit_tiny_ico2_scan_age_train": dict(
        model=dict(dim=192, depth=12, heads=3, mlp_dim=768, num_patches=320, num_vertices=153, num_channels=4,
                   num_classes=1, dim_head=64), kind="train", gflop_per_sample=13.223),
    "sit_small_ico1_birth_age_train": dict(
        model=dict(dim=384, depth=12, heads=6, mlp_dim=1536, num_patches=80, num_vertices=561, num_channels=4,
                   num_classes=1, dim_head=64), kind="train", gflop_per_sample=10.958),
    "sit_small_ico2_mpp_pretrain": dict(
        model=dict(dim=384, depth=12, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153, num_channels=4,
                   num_classes=1, dim_head=64), kind="mpp", gflop_per_sample=47.346),
    "sit_base_ico2_inference": dict(
        model=dict(dim=768, depth=12, heads=12, mlp_dim=3072, num_patches=320, num_vertices=153, num_channels=4,
                   num_classes=1, dim_head=64), kind="infer", gflop_per_sample=58.627),
}
DEFAULT_WORKLOAD = "sit_small_ico2_scan_age_train"
# `ncu --set full` summary of the kernels of the current build (scripts/ncu_top.py + scripts/ncu_summary.py); re-captured
# whenever a kernel changes -- roofline.traffic is read from it
NCU_SUMMARY = "r02c_ncu_top_summary.json"
NCU_SUMMARY_ATTN_BWD = "r02i_ncu_attn_bwd_summary.json"   # attention backward re-captured after the dead-pair shortcut


def executed_gflop(wl):
    """GFLOP per sample the engine actually owes the result.  SURVEY 8(d)'s figure (wl["gflop_per_sample"]) counts every
    token row of every block; under cls pooling (models/sit.py:78, all SiT workloads here) the last block's attention,
    to_out and FeedForward are only needed -- and only computed, engine.cu cls_last -- for token 0, so the T - 1 other
    rows of that block are subtracted: forward 4 T 64 H + 2 I D + 4 D mlp flop per row, x 3.5 / 3 / 3 in training
    (attention backward is 2.5 x its forward, a Linear's backward 2 x).  MPP decodes every token: nothing is dropped."""
    m, kind = wl["model"], wl["kind"]
    if kind == "mpp" or os.environ.get("SVIT_FULL_LAST_LAYER", "0") not in ("", "0"):
        return wl["gflop_per_sample"]
    T, D, I, mlp = m["num_patches"] + 1, m["dim"], m["heads"] * m["dim_head"], m["mlp_dim"]
    attn, lin = 4.0 * T * 64 * m["heads"], 2.0 * I * D + 4.0 * D * mlp
    dead = (T - 1) * ((3.5 * attn + 3.0 * lin) if kind == "train" else (attn + lin))
    return wl["gflop_per_sample"] - dead / 1e9


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops_burst=p["bf16_tflops"], tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm_gbs=p["hbm_gbs"], source="MEASURED_PEAKS.json")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs.

    NVML in-process (the library nvidia-smi itself reads: clocks.sm = NVML_CLOCK_SM, clocks_event_reasons.* = the bits of
    nvmlDeviceGetCurrentClocksEventReasons) every 10 ms, so that a 0.4 s timed region yields tens of samples; the
    `nvidia-smi -lms 100` subprocess of the profiling recipe is the fallback (its first line arrives after ~0.3 s)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []      # (sm_mhz, set(reasons))
        self.max_mhz = None
        self.proc = None
        self.thread = None
        self.source = None
        self._stop = threading.Event()
        self._nvml = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:  # the torch device may be remapped by CUDA_VISIBLE_DEVICES: find it by UUID
            import torch
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        return pynvml, h

    def start(self):
        try:
            self._nvml = self._nvml_handle()
            self.source = "nvml, 10 ms period"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.source = "nvidia-smi -lms 100"
        self.thread = threading.Thread(target=self._read_smi, daemon=True)
        self.thread.start()

    def _poll_nvml(self):
        pynvml, h = self._nvml
        bits = ((pynvml.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (pynvml.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (pynvml.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
        while not self._stop.is_set():
            try:
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.samples.append((mhz, {n for b, n in bits if mask & b}))
            except Exception:
                pass
            self._stop.wait(0.010)

    def _read_smi(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                try:
                    mhz = float(parts[0])
                    self.max_mhz = float(parts[1])
                except ValueError:
                    continue
                self.samples.append((mhz, {n for n, v in zip(self.NAMES, parts[3:7]) if v.lower().startswith("active")}))

    def stop(self):
        if self.thread is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvml and nvidia-smi unavailable"])
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.thread.join(timeout=5)
        sm = sorted(s[0] for s in self.samples)
        reasons = set()
        for s in self.samples:
            reasons |= s[1]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=self.max_mhz, reasons=sorted(reasons),
                    samples=len(sm), source=self.source)


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_rate(wl, batch, steps, warmup):
    """fp32 PyTorch restatement of the reference (oracle port) on the host cores: full train step / eval forward."""
    import torch
    from oracle.sit_oracle import OracleMPP, OracleSiT
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    m = wl["model"]
    model = OracleSiT(dim=m["dim"], depth=m["depth"], heads=m["heads"], mlp_dim=m["mlp_dim"],
                      num_patches=m["num_patches"], num_classes=m["num_classes"], num_channels=m["num_channels"],
                      num_vertices=m["num_vertices"], dim_head=m["dim_head"])
    x = torch.randn(batch, m["num_channels"], m["num_patches"], m["num_vertices"])
    y = torch.rand(batch) * 19 + 26
    kind = wl["kind"]
    if kind == "mpp":
        K = m["num_channels"] * m["num_vertices"]
        ssl = OracleMPP(model, m["dim"], K, "cpu", 0.5, 0.8, 0.02, m["num_channels"], m["num_vertices"])
        opt = torch.optim.AdamW(model.parameters(), lr=3e-4, weight_decay=0.0)
    elif kind == "train":
        opt = torch.optim.AdamW(model.parameters(), lr=1e-5, weight_decay=0.0)

    def step():
        if kind == "infer":
            with torch.no_grad():
                return model(x)
        opt.zero_grad()
        if kind == "mpp":
            loss, _ = ssl(x)
        else:
            loss = torch.nn.functional.mse_loss(model(x).squeeze(), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, cores


def run_reference_arm(args, wl, wl_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch
    rate, ms, cores = cpu_reference_step_rate(wl, batch, args.steps, args.warmup)
    sample = f"oracle port of models/sit.py + vit_pytorch shim, fp32, batch {batch} per step, {args.steps} steps"
    line = dict(impl="reference", metric="SiT train samples/sec", value=rate, unit="samples/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=wl_name, batch_per_step=batch, device="host cpu", **wl["model"]),
                cpu_baseline=dict(value=rate, unit="samples/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=rate, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="the reference cannot be installed (its encoder is the un-pinned third-party vit-pytorch, absent "
                     "offline; /root/reference is not on the GPU box): this arm runs the oracle port on the host cores")
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- other BASELINE configs
SWEEP_POINTS = [("sit_tiny_ico2_scan_age_train", 256), ("sit_small_ico1_birth_age_train", 256),
                ("sit_small_ico2_mpp_pretrain", 256)] + \
               [("sit_base_ico2_inference", b) for b in (64, 128, 256, 512, 1024, 2048, 4096)]


def run_sweep(args):
    """BASELINE.json configs[0,2,3,4] on one GPU: train samples/s for C1 / C3 / C4 and the C5 inference batch sweep
    (tools/testing.py:76-88 is the reference's inference path).  One record per point, written as a JSON list."""
    import torch
    import surface_vision_transformers_b200 as svit
    from surface_vision_transformers_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    peaks = load_peaks()
    records = []
    for wl_name, B in SWEEP_POINTS:
        wl = WORKLOADS[wl_name]
        m, kind = wl["model"], wl["kind"]
        torch.manual_seed(0)
        model = svit.SiT(dim=m["dim"], depth=m["depth"], heads=m["heads"], mlp_dim=m["mlp_dim"], num_patches=m["num_patches"],
                         num_classes=m["num_classes"], num_channels=m["num_channels"], num_vertices=m["num_vertices"],
                         dim_head=m["dim_head"]).to(dev)
        runner, opt = model, None
        if kind == "mpp":
            K = m["num_channels"] * m["num_vertices"]
            runner = svit.masked_patch_pretraining(transformer=model, dim_in=m["dim"], dim_out=K, device=dev, mask_prob=0.5,
                                                   replace_prob=0.8, swap_prob=0.02, channels=m["num_channels"],
                                                   num_vertices=m["num_vertices"]).to(dev)
        if kind != "infer":
            opt = svit.FusedAdamW(model.parameters(), lr=1e-5 if kind == "train" else 3e-4, weight_decay=0.0)
        else:
            model.eval()
        x = torch.randn(B, m["num_channels"], m["num_patches"], m["num_vertices"], device=dev)
        y = torch.rand(B, device=dev) * 19 + 26

        def step():
            if kind == "infer":
                with torch.no_grad():
                    return model(x)
            opt.zero_grad(set_to_none=True)
            loss = runner(x)[0] if kind == "mpp" else svit.regression_loss(runner(x), y)
            loss.backward()
            opt.step()
            return loss

        for _ in range(max(3, args.warmup)):
            step()
        torch.cuda.synchronize()
        sampler = ClockSampler(0)
        sampler.start()
        l0 = lib.svit_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        clocks = sampler.stop()
        value = B / (ms * 1e-3)
        tf = value * executed_gflop(wl) / 1e3
        rec = dict(metric="SiT train samples/sec" if kind != "infer" else "SiT inference samples/sec", value=value,
                   unit="samples/s", n_gpus=1, steps=args.steps, warmup=max(3, args.warmup), ms_per_step=ms, dtype="bf16",
                   data="synthetic", config=dict(workload=wl_name, batch_per_gpu=B, **m), clocks=clocks,
                   gpu_launches=int(lib.svit_launch_count() - l0),
                   roofline=dict(bound="tensor", scope="whole step", achieved=tf, unit="TFLOP/s",
                                 peak=peaks["tflops_sustained"], frac=tf / peaks["tflops_sustained"],
                                 gflop_per_sample=executed_gflop(wl), gflop_per_sample_all_rows=wl["gflop_per_sample"],
                                 peak_source=peaks["source"] + " (sustained)"))
        records.append(rec)
        print(json.dumps(rec), flush=True)
        del model, runner, opt, x, y
        torch.cuda.empty_cache()
    with open(args.sweep, "w") as f:
        json.dump(records, f, indent=1)


# --------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=16)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-input", default="fp32", choices=["bf16", "fp32"],
                    help="dtype of the pinned host batches of the e2e leg: fp32 (what the reference's loader yields, default) or "
                         "bf16 (rounded once on the host; the patch-embedding GEMM consumes bf16 either way, results are "
                         "bit-identical, half the H2D bytes -- measured on one GPU: no faster, the copy is already hidden)")
    ap.add_argument("--sweep", metavar="OUT.json", default=None,
                    help="instead of the headline line: time the other BASELINE.json configs (C1 tiny / C3 ico-1 / C4 MPP "
                         "training at the per-GPU batch, C5 SiT-base inference at batch 64..4096) and write one record per "
                         "point to OUT.json (device-resident timing, same clocks / roofline fields)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    wl_name = args.workload
    wl = WORKLOADS[wl_name]
    if args.impl == "reference":
        run_reference_arm(args, wl, wl_name)
        return
    if args.sweep:
        run_sweep(args)
        return

    import torch
    import torch.distributed as dist
    import surface_vision_transformers_b200 as svit
    from surface_vision_transformers_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL's kernels on a high-priority stream: at a communication window (ddp.py) its few CTAs are placed before the
        # pending CTAs of the attention backward they share the chip with (SVIT_NCCL_HIGH_PRIO=0 for A/B)
        opts = None
        if os.environ.get("SVIT_NCCL_HIGH_PRIO", "1") != "0":
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    lib = _lib.load()
    peaks = load_peaks()

    m = wl["model"]
    kind = wl["kind"]
    B = args.batch
    cls_only_last = executed_gflop(wl) != wl["gflop_per_sample"]
    torch.manual_seed(0)
    model = svit.SiT(dim=m["dim"], depth=m["depth"], heads=m["heads"], mlp_dim=m["mlp_dim"], num_patches=m["num_patches"],
                     num_classes=m["num_classes"], num_channels=m["num_channels"], num_vertices=m["num_vertices"],
                     dim_head=m["dim_head"]).to(dev)
    runner = model
    if kind == "mpp":
        K = m["num_channels"] * m["num_vertices"]
        runner = svit.masked_patch_pretraining(transformer=model, dim_in=m["dim"], dim_out=K, device=dev, mask_prob=0.5,
                                               replace_prob=0.8, swap_prob=0.02, channels=m["num_channels"],
                                               num_vertices=m["num_vertices"]).to(dev)
    if world > 1:
        runner = svit.DataParallel(runner)
    opt = svit.FusedAdamW(model.parameters(), lr=1e-5 if kind == "train" else 3e-4, weight_decay=0.0) \
        if kind != "infer" else None

    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    x_dev = torch.randn(B, m["num_channels"], m["num_patches"], m["num_vertices"], device=dev, generator=g)
    y_dev = torch.rand(B, device=dev, generator=g) * 19 + 26
    # e2e batches: pinned host memory; bf16 staging halves the H2D bytes (MPP keeps fp32: its loss compares with the input)
    e2e_bf16 = args.e2e_input == "bf16" and kind != "mpp"
    x_host = (x_dev.bfloat16() if e2e_bf16 else x_dev).cpu().pin_memory()
    y_host = y_dev.cpu().pin_memory()

    def step(x, y):
        if kind == "infer":
            with torch.no_grad():
                return model(x).sum()
        opt.zero_grad(set_to_none=True)
        if kind == "mpp":
            loss, _ = runner(x)
        else:
            loss = svit.regression_loss(runner(x), y)     # MSELoss(mean) of train.py:245-248 in one launch
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- warm-up ----
    for _ in range(args.warmup):
        step(x_dev, y_dev)
    torch.cuda.synchronize()

    # ---- device-resident timed region ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = lib.svit_launch_count()
    ms_total = timed(lambda: step(x_dev, y_dev), args.steps)
    launches = lib.svit_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end-to-end: host inputs in, loss out, every step ----
    # Every step's batch starts in pinned HOST memory and is copied inside the timed region; every step's loss is
    # read back on the host inside the timed region.  The public API a user would write this loop with is
    # svit.DevicePrefetcher (copy of batch i+1 on a side stream under the kernels of batch i); the loss of step i is
    # read while step i+1 is already enqueued (one pinned scalar per step, two in flight).
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    losses = []

    def e2e_run(steps):
        pending = None
        batches = ((x_host, y_host) for _ in range(steps))
        for i, (xs, ys) in enumerate(svit.DevicePrefetcher(batches, dev)):
            loss = step(xs, ys)
            buf = loss_host[i & 1]
            buf.copy_(loss.detach().reshape(()), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending[1].synchronize()
                losses.append(float(pending[0]))
            pending = (buf, ev)
        pending[1].synchronize()
        losses.append(float(pending[0]))

    e2e_run(3)
    losses.clear()
    ms_e2e = timed(lambda: e2e_run(args.steps), 1)
    assert len(losses) == args.steps
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    h2d = x_host.numel() * x_host.element_size() + y_host.numel() * 4
    d2h = 4

    # ---- dominant kernel (MLP up-projection GEMM + bias + GELU epilogue) timed alone ----
    roof = None
    if rank == 0:
        M = B * (m["num_patches"] + 1)
        D, H4 = m["dim"], m["mlp_dim"]
        A = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
        W = (torch.randn(H4, D, device=dev) * 0.05).bfloat16()
        bias = torch.zeros(H4, device=dev)
        o1 = torch.empty(M, H4, device=dev, dtype=torch.bfloat16)
        o2 = torch.empty(M, H4, device=dev, dtype=torch.bfloat16)
        mode = 4 if kind == "infer" else 5   # training stores gelu'(u) and gelu(u) (EPI_GELU_GRAD), inference gelu(u) only
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        st = _lib.vp(torch.cuda.current_stream().cuda_stream)

        def gemm():
            _lib.check(lib.svit_gemm_tn(_lib.ptr(A), _lib.ptr(W), _lib.ptr(o1), _lib.ptr(o2), _lib.vp(0), _lib.ptr(bias),
                                        _lib.vp(0), 1, M, H4, D, D, D, H4, mode, 0, sms, st), "gemm")
        # `peak` below is the BURST figure of MEASURED_PEAKS.json (a kernel timed alone), so the kernel is timed in burst
        # conditions too: one idle second first -- right after the long power-capped step the SM clock is still at its
        # sustained value (~1.7 GHz) and a 20-launch loop (3-6 ms) is over before it recovers.  The clock is sampled
        # through NVML right after the loop and reported next to the duration.
        def sm_clock_now():
            try:
                import pynvml
                pynvml.nvmlInit()
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(dev).uuid)).encode())
                except Exception:
                    h = pynvml.nvmlDeviceGetHandleByIndex(dev.index or 0)
                return float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            except Exception:
                return None

        torch.cuda.synchronize()
        time.sleep(1.0)
        for _ in range(3):
            gemm()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_it = 20
        e0.record()
        for _ in range(n_it):
            gemm()
        e1.record()
        k_mhz = sm_clock_now()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / n_it
        flops = 2.0 * M * H4 * D
        ach = flops / (k_ms * 1e-3) / 1e12
        step_tf = value / world * executed_gflop(wl) / 1e3
        # DRAM bytes of this kernel from the committed `ncu --set full` capture (profiles/), same shape only
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", NCU_SUMMARY)) as f:
                prof = json.load(f)["gemm_tn_kernel<__nv_bfloat16, 5, 192, 2, 1>"]
            if kind != "infer" and (M, H4, D) == (82176, 1536, 384):
                traffic = prof["dram_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            pass
        roof_gemm = dict(bound="tensor", kernel="gemm_tn_kernel<bf16,%s,192,cta_group::2> (fc1 + bias + exact GELU%s, M=%d N=%d K=%d)" % ("EPI_GELU_ONLY" if mode == 4 else "EPI_GELU_GRAD", "" if mode == 4 else " and its derivative", M, H4, D),
                    achieved=ach, peak=peaks["tflops_burst"], unit="TFLOP/s", frac=ach / peaks["tflops_burst"],
                    traffic=traffic, traffic_source="profiles/" + NCU_SUMMARY + " (dram__bytes_read.sum + "
                    "dram__bytes_write.sum, one ncu --set full launch)" if traffic else None,
                    algorithmic_bytes=2.0 * M * D + 2.0 * H4 * D + (2 if mode == 5 else 1) * 2.0 * M * H4,
                    peak_source=peaks["source"] + " (burst: kernel timed alone)",
                    us_per_launch=k_ms * 1e3, flops_per_launch=flops, sm_mhz_during_launches=k_mhz)
        roof_gemm["hbm_view"] = dict(achieved=roof_gemm["algorithmic_bytes"] / (k_ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"],
                                     unit="GB/s", frac=roof_gemm["algorithmic_bytes"] / (k_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                     note="arithmetic intensity %.0f flop/B is below the machine balance: by the roofline "
                                          "model this launch is HBM-bound" % (flops / roof_gemm["algorithmic_bytes"]))

        # ---- the kernel with the largest share of the step: fused attention (backward when training) ----
        Hh, T = m["heads"], m["num_patches"] + 1
        inner = Hh * 64
        qkv = torch.randn(B, T, 3 * inner, device=dev).bfloat16()
        o = torch.empty(B, T, inner, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, Hh, T, device=dev)
        scale = ctypes.c_float(0.125)
        _lib.check(lib.svit_attn_fwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(lse), B, Hh, T, scale, st), "attn_fwd")
        if kind == "infer":
            def attn():
                _lib.check(lib.svit_attn_fwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(lse), B, Hh, T, scale, st), "attn_fwd")
            a_flops, a_name, a_key = 4.0 * T * T * 64 * B * Hh, "attn_fwd_kernel", "attn_fwd"
            a_bytes = 2.0 * B * T * (3 * inner + inner) + 4.0 * B * Hh * T
        else:
            do = torch.randn(B, T, inner, device=dev).bfloat16()
            dqkv = torch.empty_like(qkv)

            def attn():
                _lib.check(lib.svit_attn_bwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(do), _lib.ptr(lse), _lib.ptr(dqkv), B, Hh, T, scale, st), "attn_bwd")
            a_flops, a_name, a_key = 10.0 * T * T * 64 * B * Hh, "attn_bwd_kernel", "attn_bwd"
            a_bytes = 2.0 * B * T * (3 * inner + inner + inner + 3 * inner) + 4.0 * B * Hh * T
        torch.cuda.synchronize()
        time.sleep(1.0)
        for _ in range(3):
            attn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n_it):
            attn()
        e1.record()
        a_mhz = sm_clock_now()
        torch.cuda.synchronize()
        a_ms = e0.elapsed_time(e1) / n_it
        a_ach = a_flops / (a_ms * 1e-3) / 1e12
        a_traffic = None
        try:
            a_src = NCU_SUMMARY_ATTN_BWD if (a_key == "attn_bwd" and os.path.exists(
                os.path.join(ROOT, "profiles", NCU_SUMMARY_ATTN_BWD))) else NCU_SUMMARY
            with open(os.path.join(ROOT, "profiles", a_src)) as f:
                prof = json.load(f)[a_key + "_kernel"]
            if (B, Hh, T) == (256, 6, 321):
                a_traffic = prof["dram_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            a_src = NCU_SUMMARY
        roof = dict(bound="tensor", kernel="%s (one CTA per (sample, head); B=%d H=%d T=%d d=64; %s*T^2*64 flop per head at the "
                                          "unpadded T)" % (a_name, B, Hh, T, "4" if kind == "infer" else "10"),
                    achieved=a_ach, peak=peaks["tflops_burst"], unit="TFLOP/s", frac=a_ach / peaks["tflops_burst"],
                    traffic=a_traffic, traffic_source="profiles/" + a_src + " (dram__bytes_read.sum + "
                    "dram__bytes_write.sum, one ncu --set full launch)" if a_traffic else None,
                    algorithmic_bytes=a_bytes, peak_source=peaks["source"] + " (burst: kernel timed alone)",
                    us_per_launch=a_ms * 1e3, flops_per_launch=a_flops, sm_mhz_during_launches=a_mhz,
                    share_of_step=a_ms * (m["depth"] - (1 if cls_only_last else 0)) / ms_step,
                    note="fraction of the dense bf16 tensor peak at the algorithmic flop count; T=%d pads to 128x96 tiles "
                         "(executed MMA work is 1.4x the algorithmic count at T=321) and the kernel is bound by the latency "
                         "chain of its compute warps through one P tile per step, not by the tensor pipe (DESIGN.md 3, 3b)" % T,
                    step_achieved_tflops=step_tf, step_frac_of_sustained=step_tf / peaks["tflops_sustained"],
                    step_frac_of_nominal_2250=step_tf / 2250.0,
                    step_gflop_per_sample=executed_gflop(wl), step_gflop_per_sample_all_rows=wl["gflop_per_sample"],
                    step_flop_note="whole-step TFLOP/s counts the flops the result needs: with cls pooling the last "
                                   "block's attention / to_out / FeedForward run for token 0 only (exact, engine.cu "
                                   "cls_last), so its other T-1 rows are not counted; SURVEY 8(d)'s all-rows figure is "
                                   "step_gflop_per_sample_all_rows" if cls_only_last else None)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        rate, ms_cpu, cores = cpu_reference_step_rate(wl, args.cpu_batch, 10, 3)   # 3 warm-up + >= 10 timed
        cpu = dict(value=rate, unit="samples/s", cores=cores, kind="port",
                   sample=f"oracle port (fp32 PyTorch restatement of the reference), batch {args.cpu_batch}, 3 warm-up + 10 "
                          f"timed steps of the same workload ({ms_cpu:.0f} ms/step)")

    if rank == 0:
        line = dict(metric="SiT train samples/sec" if kind != "infer" else "SiT inference samples/sec", value=value,
                    unit="samples/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload=wl_name, batch_per_gpu=B, global_batch=B * world,
                                parallelism=f"dp{world}" if world > 1 else "single", optimizer="FusedAdamW",
                                l2_policy="inputs and activations larger than L2 (batch %.0f MB, activations > 10 GB)" % (x_dev.numel() * 4 / 1e6),
                                last_block="token 0 only (cls pooling: the head reads x[:, 0]; same outputs and gradients)"
                                if cls_only_last else "all rows", **m),
                    clocks=clocks, e2e=dict(value=e2e_value, unit="samples/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                                            ms_per_step=ms_e2e / args.steps,
                                            input="pinned host batch as %s%s" % (
                                                "bf16" if e2e_bf16 else "fp32",
                                                " (rounded once on the host: bit-identical results, the embedding GEMM "
                                                "consumes bf16 either way)"
                                                if e2e_bf16 else "")),
                    gpu_launches=int(launches), roofline=roof, roofline_gemm=roof_gemm if rank == 0 else None, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
