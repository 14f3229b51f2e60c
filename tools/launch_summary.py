"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches / total us / share,
for the LAST `1/passes` of the launches (the final pass of tools/engine_pass.py).
usage: python tools/launch_summary.py launches.csv [passes] > profiles/xxx.txt"""
import csv, re, sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
data = [dict(zip(hdr, r)) for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("Metric Name")] == "gpu__time_duration.sum"]
n = len(data) // passes
last = data[len(data) - n:]
agg = {}
for d in last:
    name = re.sub(r"\(.*", "", d["Kernel Name"])[:110]
    v = float(d["Metric Value"].replace(",", ""))
    if d["Metric Unit"].startswith("ns"):
        v /= 1e3
    elif d["Metric Unit"].startswith("ms"):
        v *= 1e3
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += v
tot = sum(v for _, v in agg.values())
print(f"# {sys.argv[1]}: last pass of {passes}: {n} launches, {tot / 1e3:.3f} ms of kernel time (cold-cache, serialised under ncu)")
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f}% x{cnt:4d}  {name}")
