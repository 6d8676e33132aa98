"""-m "not gpu": the parts of bench.py that do not need a GPU -- the reference arm's JSON line (the driver parses it), the
flop accounting behind `roofline.step_*`, and the clock sampler's behaviour on a box without NVML / nvidia-smi."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` runs the oracle port on the host cores (no GPU, no product code) and prints ONE JSON line with
    the keys of the bench contract, `impl: reference`, a `cpu_baseline` describing the run and an `e2e` that repeats the value."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "sit_tiny_ico2_scan_age_train"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["config"]["workload"] == "sit_tiny_ico2_scan_age_train" and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_flop_accounting_matches_survey_8d():
    """2 MACs of every matmul at the unpadded T, training = 3 x forward - patch-embedding dgrad.  The executed
    count drops the T - 1 dead rows of the last block under cls pooling and nothing for MPP."""
    wls = bench.WORKLOADS if isinstance(bench.WORKLOADS, dict) else {w["name"]: w for w in bench.WORKLOADS}
    wl = wls["sit_small_ico2_scan_age_train"]
    m = wl["model"]
    T, D, I, mlp, N = m["num_patches"] + 1, m["dim"], m["heads"] * 64, m["mlp_dim"], m["num_patches"]
    K = m["num_channels"] * m["num_vertices"]
    pe = 2.0 * N * K * D
    layer = 2.0 * T * D * 3 * I + 4.0 * m["heads"] * T * T * 64 + 2.0 * T * I * D + 4.0 * T * D * mlp
    fwd = pe + m["depth"] * layer + 2.0 * D
    train = 3.0 * fwd - pe
    assert abs(train / 1e9 - wl["gflop_per_sample"]) < 0.01            # 46.895
    dead_row = 3.5 * 4.0 * T * 64 * m["heads"] + 3.0 * (2.0 * I * D + 4.0 * D * mlp)
    assert abs(bench.executed_gflop(wl) - (train - (T - 1) * dead_row) / 1e9) < 1e-3   # 43.795 (from the rounded 46.895)
    mpp = wls["sit_small_ico2_mpp_pretrain"]
    assert bench.executed_gflop(mpp) == mpp["gflop_per_sample"]


def test_clock_sampler_without_a_gpu_reports_unavailable():
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    if out["sm_mhz"] is None:                                           # this container: no NVML, no nvidia-smi
        assert out["reasons"] and "unavailable" in out["reasons"][0]
