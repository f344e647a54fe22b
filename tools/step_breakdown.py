"""Per-kernel (name, grid) breakdown of ONE steady-state train step from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv`): the launches between two consecutive G-net Adam updates."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
adam = [i for i, r in enumerate(rows) if "adam" in r["Kernel Name"]]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 2          # step index (0-based) to show
s, e = adam[3 * which - 1] + 1 if which else 0, adam[3 * which + 2] + 1
st = rows[s:e]


def us(r):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    return v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3


tot = sum(us(r) for r in st)
print("step %d: %d launches, %.1f us of kernel time" % (which, len(st), tot))
agg = collections.defaultdict(lambda: [0, 0.0])
byname = collections.defaultdict(lambda: [0, 0.0])
for r in st:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    agg[(name, r["Grid Size"])][0] += 1
    agg[(name, r["Grid Size"])][1] += us(r)
    byname[name][0] += 1
    byname[name][1] += us(r)
print("--- by kernel")
for k, (c, t) in sorted(byname.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%5.1f%% %9.1f us %4d  %s" % (100 * t / tot, t, c, k[:90]))
print("--- by kernel and grid")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%9.1f us %4d  avg %7.1f  %s grid=%s" % (t, c, t / c, k[0][:60], k[1]))
