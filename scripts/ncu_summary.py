"""Condenses `ncu --set full` reports into one JSON (profiles/rNN_ncu_full_summary.json): per kernel launch the duration,
DRAM bytes (read + write = `roofline.traffic`), tensor / issue / XU / FMA / ALU pipe utilisation, registers, occupancy.

    python scripts/ncu_summary.py gpurun_out/a.ncu-rep gpurun_out/b.ncu-rep ... -o profiles/r02_ncu_full_summary.json
"""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.max",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield {h: (v, u) for h, v, u in zip(hdr, r, units)}


def main():
    args = sys.argv[1:]
    out = args[args.index("-o") + 1] if "-o" in args else None
    reps = [a for a in args if a.endswith(".ncu-rep")]
    summary = {}
    for rep in reps:
        for r in rows_of(rep):
            name = r["Kernel Name"][0]
            key = name.split("(")[0].replace("void ", "").replace("svit::", "")
            n, k = 1, key
            while k in summary:
                n += 1
                k = f"{key}#{n}"
            rec = {"kernel": name, "report": rep}
            for m in KEYS:
                if m in r:
                    v, u = r[m]
                    try:
                        rec[m] = float(v.replace(",", ""))
                    except ValueError:
                        rec[m] = v
                    if u:
                        rec[m + ".unit"] = u
            def b(m):
                if m not in r:
                    return None
                v, u = r[m]
                return float(v.replace(",", "")) * UNIT.get(u, 1.0)
            rd, wr = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
            if rd is not None and wr is not None:
                rec["dram_bytes_per_launch"] = rd + wr
                t, tu = r["gpu__time_duration.sum"]
                us = float(t.replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(tu, 1.0)
                rec["us"] = us
                rec["dram_gb_per_s"] = (rd + wr) / us / 1e3
            summary[k] = rec
    txt = json.dumps(summary, indent=1)
    if out:
        open(out, "w").write(txt)
    else:
        print(txt)


if __name__ == "__main__":
    main()
