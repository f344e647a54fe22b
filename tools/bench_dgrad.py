"""Graph-timed device time of the trunk conv forward and its stride-1 input gradient (with / without border-strip regions)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa
from vts_b200 import ops
from tools.bench_conv import timeit

for s in (128, 192):
    x = torch.randn(1, s, s, 256, device="cuda")
    w = torch.randn(256, 256, 3, 3, device="cuda") / math.sqrt(2304)
    _, op = ops.norm_act_pad(x, pad=1, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2)
    _, dop = ops.norm_act_pad(x, pad=2, pad_mode=ops.PAD_ZERO, fmt=ops.FMT_BF16X2)
    pk0 = ops.PackedWeights(w, 0, want_f32=False, want_bf16=True)
    pk1 = ops.PackedWeights(w, 1, want_f32=False, want_bf16=True)
    y = torch.empty(1, s, s, 256, device="cuda")
    t = timeit(lambda: ops.conv2d_fwd(op, pk0, 1, 0, s, s, stats_mode=ops.NORM_NONE, impl=ops.IMPL_TC, out=y))
    print("%dx%d fwd (no stats): %.1f us" % (s, s, t * 1e3))
    t = timeit(lambda: ops.conv2d_dgrad_s1(dop, pk1))
    a = ops.conv2d_dgrad_s1(dop, pk1)
    b, _ = ops.conv2d_fwd(dop, pk1, 1, 0, s + 2, s + 2, impl=ops.IMPL_TC)
    torch.cuda.synchronize()
    print("   dgrad_s1 (regions): %.1f us   max|diff| vs one-region launch: %.3e" % (t * 1e3, (a - b).abs().max().item()))
    t = timeit(lambda: ops.conv2d_fwd(dop, pk1, 1, 0, s + 2, s + 2, impl=ops.IMPL_TC))
    print("   dgrad as one region: %.1f us" % (t * 1e3))
