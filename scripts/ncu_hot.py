"""Lists the hottest SASS instructions (by warp-stall samples) of an ncu report's source page."""
import csv, subprocess, sys
# usage: ncu_hot.py report.ncu-rep [top N] [kernel regex] [invocation nr]   (the last two pick one launch of a multi-kernel report)
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sel = ["--kernel-id", f"::regex:{sys.argv[3]}:{sys.argv[4] if len(sys.argv) > 4 else 1}"] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[hi + 1:]):
    try:
        n = int(r[ci["# Samples"]])
    except Exception:
        continue
    data.append((n, idx, r))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, idx, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{n:6d} {100*n/tot:5.1f}%  #{idx:5d} {r[ci['Source']].strip()[:90]:90s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
