"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name launches, mean / total device time, share."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
    key = (name, r["Grid Size"], r["Block Size"])
    agg.setdefault(key, []).append(float(r["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{name[:44]:44s} grid {grid:14s} x{len(v):4d}  mean {sum(v)/len(v)/1e3:9.1f} us  total {sum(v)/1e6:8.2f} ms  {100*sum(v)/tot:5.1f} %")
