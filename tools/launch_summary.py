"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (markdown table on stdout)."""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in r:
    if len(row) < len(hdr) or row[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = row[ix["Kernel Name"]].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    m = re.search(r"(sim::\w+)", name)
    key = m.group(1) if m else re.sub(r"<.*", "", name)[:60]
    if "gemm_split3" in key:
        key += " grid=" + row[ix["Grid Size"]].split(",")[0].strip("( ")
    v = float(row[ix["Metric Value"]])
    u = row[ix["Metric Unit"]]
    v = v / 1000.0 if u in ("nsecond", "ns") else v * 1000.0 if u in ("msecond", "ms") else v
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches\n")
print("| kernel | launches | total us | us / launch | share |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
