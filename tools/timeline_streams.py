"""Per-stream kernel sequence of one replayed train step (torch.profiler chrome trace): the launching stream's chain with
durations and gaps.   python tools/timeline_streams.py [size] [top]"""
import collections, json, os, sys, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa
from oracle import skit_oracle as O  # synthetic batch factory only

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m = vts_b200.SinSKITGModel(vts_b200.default_options())
m.set_input(O.synthetic_batch(size, NT=64, seed=0))
for _ in range(5):
    m.optimize_parameters(1)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.optimize_parameters(1)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
by = collections.defaultdict(list)
for e in ev:
    by[e["args"].get("stream")].append(e)
print("streams:", {k: (len(v), round(sum(x["dur"] for x in v) / 1e3, 2)) for k, v in by.items()})
main = max(by.items(), key=lambda kv: sum(x["dur"] for x in kv[1]))[0]
seq = by[main]
print("main stream %s: %d kernels, busy %.2f ms, span %.2f ms" % (main, len(seq), sum(x["dur"] for x in seq) / 1e3, (seq[-1]["ts"] + seq[-1]["dur"] - seq[0]["ts"]) / 1e3))
# chain listing, merged by consecutive name
gaps = []
prev_end = seq[0]["ts"]
rows = []
for e in seq:
    gap = e["ts"] - prev_end
    gaps.append(gap)
    rows.append((e["ts"] - t0, e["dur"], gap, e["name"][:70]))
    prev_end = e["ts"] + e["dur"]
print("total gap on main stream: %.2f ms; gaps > 20us:" % (sum(g for g in gaps if g > 0) / 1e3))
for r in rows:
    if r[2] > 20:
        print("   t=%8.1f us  gap %7.1f us before %s" % (r[0], r[2], r[3]))
agg = collections.Counter()
for r in rows:
    agg[r[3][:55]] += r[1]
print("main-stream kernel time by name:")
for k, v in agg.most_common(16):
    print("   %8.1f us  %s" % (v, k))
