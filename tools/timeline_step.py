"""Kernel timeline of one replayed train step (torch.profiler / CUPTI): per-stream busy time, union busy time, idle gaps,
and the kernels on the launching (critical) stream.   python tools/timeline_step.py [size]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa: E402
from oracle import skit_oracle as O  # noqa: E402  (synthetic batch factory only)

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
arch = sys.argv[2] if len(sys.argv) > 2 else "B"
if arch == "B":
    opt = vts_b200.default_options()
elif arch == "N":      # arch B + PatchNCE (BASELINE.json configs[2] when size = 768)
    opt = vts_b200.default_options(lambda_NCE=1.0)
elif arch == "L":      # arch B + the reference's default LPIPS terms
    opt = vts_b200.default_options(lambda_G1_lpips=1.0, lambda_G2_lpips=10.0, allow_random_lpips=True)
else:
    opt = vts_b200.default_options(netG="unet256_custom", ngf=10, ndf=8)
torch.manual_seed(0)
m = vts_b200.SinSKITGModel(opt)
m.set_input(O.synthetic_batch(size, NT=64, seed=0))
for _ in range(5):
    m.optimize_parameters(1)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.optimize_parameters(1)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ks = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "stream", None) if hasattr(e, "stream") else None) for e in evs))
if not ks:
    print("no kernel events")
    sys.exit(0)
t0, t1 = ks[0][0], max(k[1] for k in ks)
print("kernels: %d  span: %.2f ms" % (len(ks), (t1 - t0) / 1e3))
# union busy time
busy, cur_s, cur_e = 0.0, None, None
for s, e, _, _ in ks:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print("union busy: %.2f ms  (idle %.2f ms)   sum of kernel durations: %.2f ms" % (busy / 1e3, (t1 - t0 - busy) / 1e3, sum(e - s for s, e, _, _ in ks) / 1e3))
# concurrency histogram: time with k kernels in flight
pts = []
for s, e, _, _ in ks:
    pts.append((s, 1)); pts.append((e, -1))
pts.sort()
hist = collections.Counter()
lvl, last = 0, pts[0][0]
for t, d in pts:
    hist[lvl] += t - last
    last = t
    lvl += d
print("time by number of kernels in flight:", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})
# phases: biggest single-kernel-in-flight stretches by kernel name
solo = collections.Counter()
active = []
import heapq
events = sorted([(s, 0, i) for i, (s, e, _, _) in enumerate(ks)] + [(e, 1, i) for i, (s, e, _, _) in enumerate(ks)])
live = set()
last = events[0][0]
for t, typ, i in events:
    if len(live) == 1:
        solo[ks[next(iter(live))][2][:60]] += t - last
    last = t
    if typ == 0:
        live.add(i)
    else:
        live.discard(i)
print("time with exactly ONE kernel in flight, by kernel:")
for name, v in solo.most_common(18):
    print("   %8.1f us  %s" % (v, name))
agg = collections.Counter()
for s, e, name, _ in ks:
    agg[name[:60]] += e - s
print("kernel time by name:")
for name, v in agg.most_common(30):
    print("   %8.1f us  %s" % (v, name))
