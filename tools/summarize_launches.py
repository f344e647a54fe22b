"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[name][0] += 1; agg[name][1] += us; tot += us
print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%7.1f%%  %10.1f us  %5d  %s" % (100 * us / tot, us, c, name[:110]))
