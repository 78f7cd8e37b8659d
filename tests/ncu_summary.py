"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
tot = 0.0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v_us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0 if unit in ("ms", "msecond") else v
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + v_us)
    tot += v_us
print(f"total {tot / 1000:.3f} ms over {sum(c for c, _ in agg.values())} launches")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1000:10.3f} ms {100 * t / tot:6.2f}%  {c:6d} launches  avg {t / c:9.1f} us  {name}")
