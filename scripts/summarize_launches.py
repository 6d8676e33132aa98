"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections, csv, re, sys

def summarize(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else v * 1e3 if r[mu] == "ms" else v
        name = re.sub(r"\(.*", "", r[kn])
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    out = [f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches"]
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{v:10.1f} us {100 * v / tot:5.1f}%  n={n:4d}  avg={v / n:8.1f} us  {k[:100]}")
    return "\n".join(out)

if __name__ == "__main__":
    print(summarize(sys.argv[1]))
