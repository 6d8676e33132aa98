"""Per-kernel SASS evidence of the shipped library: counts of the Blackwell tensor / TMEM / TMA instructions
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit),
of legacy HMMA (mma.sync; must be 0) and of local-memory accesses (LDL / STL; spills), plus registers per thread.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "surface_vision_transformers_b200", "libsvit_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+)", line)
    if m and cur:
        regs[cur] = int(m.group(1))
def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "LDL", "STL"]
rows, name, counts = [], None, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, counts))
        name, counts = m.group(1), dict.fromkeys(MN, 0)
        continue
    if name:
        for k in MN:
            if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line):
                if k == "HMMA" and "UTCHMMA" in line:
                    continue
                counts[k] += 1
if name:
    rows.append((name, counts))
print(f"SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)\n")
print(f"{'kernel':90s} {'regs':>4s} " + " ".join(f"{k:>7s}" for k in MN))
for n, c in sorted(rows, key=lambda r: -r[1]["UTCHMMA"]):
    d = demangle(n)
    d = re.sub(r"\((int|bool)\)", "", d)
    d = re.sub(r"\(.*", "", d).replace("svit::", "").replace("void ", "")
    if not (c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"] or "kernel" in d):
        continue
    print(f"{d[:90]:90s} {regs.get(n, 0):4d} " + " ".join(f"{c[k]:7d}" for k in MN))
tot = {k: sum(c[k] for _, c in rows) for k in MN}
print(f"\n{'total':90s}      " + " ".join(f"{tot[k]:7d}" for k in MN))
print("\nHMMA (mma.sync) instructions in the library:", tot["HMMA"])
