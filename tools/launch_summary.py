"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv [skip_fraction]
(skip_fraction = leading share of launches to drop, e.g. 0.5 to keep the second of two identical passes)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr, data = None, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
data = data[int(len(data) * skip):]
agg = collections.OrderedDict()
for r in data:
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1000 if r["Metric Unit"] == "ns" else v * 1000 if r["Metric Unit"] == "ms" else v
    k = (r["Kernel Name"].split("(")[0][-48:], r["Grid Size"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{len(data)} launches, {tot:.1f} us")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print(f"{k[0]:48s} {k[1]:16s} n={a[0]:4d} {a[1] / a[0]:9.1f} us each {a[1]:10.1f} us {a[1] / tot * 100:5.1f}%")
