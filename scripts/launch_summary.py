"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            k = d["Kernel Name"].split("(")[0][:48]
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            agg[k][0] += 1
            agg[k][1] += v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
tot = sum(t for _, t in agg.values())
print(f"total {tot:.3f} ms")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:9.3f} ms {100 * t / tot:5.1f} %  {c:4d} launches  {k}")
