"""Graph-timed device time of the generator head conv (64 -> 5, k7) at 512x512: fp32 head kernel vs the tcgen05 halo kernel."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa
from vts_b200 import ops
from tools.bench_conv import timeit
s = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = torch.randn(1, s, s, 64, device="cuda")
w = torch.randn(5, 64, 7, 7, device="cuda") / math.sqrt(64 * 49)
b = torch.randn(5, device="cuda")
_, op = ops.norm_act_pad(x, pad=3, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2)
pk = ops.PackedWeights(w, 0, want_f32=False, want_bf16=True)
y1, _ = ops.conv2d_fwd(op, pk, 1, 0, s, s, bias=b)
y2, _ = ops.conv2d_fwd(op, pk, 1, 0, s, s, bias=b, impl=ops.IMPL_TC)
torch.cuda.synchronize()
print("rel diff head kernel vs tcgen05:", ((y1 - y2).norm() / y2.norm()).item())
fl = 2.0 * 49 * 64 * 5 * s * s
t = timeit(lambda: ops.conv2d_fwd(op, pk, 1, 0, s, s, bias=b))
print("fp32 head kernel: %.1f us  (%.1f TFLOP/s)" % (t * 1e3, fl / t / 1e9))
t = timeit(lambda: ops.conv2d_fwd(op, pk, 1, 0, s, s, bias=b, impl=ops.IMPL_TC))
print("tcgen05 halo (N=16): %.1f us" % (t * 1e3))
