"""Text table from a profiles/rNN_ncu_full_summary.json: one line per kernel launch."""
import json, sys
d = json.load(open(sys.argv[1]))
print(f"{'kernel':62s} {'us':>8s} {'DRAM MB':>8s} {'GB/s':>7s} {'tensor%':>7s} {'issue%':>7s} {'xu%':>6s} {'regs':>5s} {'LDL':>8s}")
for k, r in d.items():
    if not (k.startswith("gemm") or "kernel" in k):
        continue
    print(f"{k[:62]:62s} {r.get('us', 0):8.1f} {r.get('dram_bytes_per_launch', 0)/1e6:8.1f} {r.get('dram_gb_per_s', 0):7.0f} "
          f"{r.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):7.1f} "
          f"{r.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):7.1f} "
          f"{r.get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 0):6.1f} "
          f"{int(r.get('launch__registers_per_thread', 0)):5d} {int(r.get('sass__inst_executed_local_loads', 0)):8d}")
