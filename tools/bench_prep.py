"""Micro-benchmark of the NHWC prep / norm-backward kernels at the trunk shape (graph-timed device time)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa
from vts_b200 import ops
from tools.bench_conv import timeit

def run(n, h, w, c):
    raw = torch.randn(n, h, w, c, device="cuda")
    dpad = torch.randn(n, h + 2, w + 2, c, device="cuda")
    dadd = torch.randn(n, h, w, c, device="cuda")
    _, st = None, torch.zeros(n, c, 2, dtype=torch.float64, device="cuda")
    st[..., 1] = h * w
    mr = ops.stats_finalize(st, h * w)
    mb = raw.numel() * 4 / 1e6
    t = timeit(lambda: ops.act_norm_bwd_reduce(raw.shape, dpad, 1, ops.PAD_REFLECT, dadd, raw, mr, ops.NORM_INSTANCE, None, None, ops.ACT_RELU))
    print("%dx%dx%dx%d  act_norm_bwd_reduce IN+relu+fold: %.1f us (%.0f GB/s over 4 streams)" % (n, h, w, c, t * 1e3, 4 * mb / t / 1e3 * 1e3 / 1e3))
    t = timeit(lambda: ops.act_norm_bwd_reduce(raw.shape, dpad, 1, ops.PAD_REFLECT, dadd, raw, None, ops.NORM_NONE, None, None, ops.ACT_RELU))
    print("   same without norm (no sums/atomics): %.1f us" % (t * 1e3))
    g, sums = ops.act_norm_bwd_reduce(raw.shape, dpad, 1, ops.PAD_REFLECT, dadd, raw, mr, ops.NORM_INSTANCE, None, None, ops.ACT_RELU)
    t = timeit(lambda: ops.norm_bwd_apply(g, raw, mr, ops.NORM_INSTANCE, None, sums, h * w, pad=2, fmt=ops.FMT_BF16X2))
    print("   norm_bwd_apply -> bf16x2 pad 2: %.1f us" % (t * 1e3))
    t = timeit(lambda: ops.norm_act_pad(raw, mr, ops.NORM_INSTANCE, act=ops.ACT_RELU, pad=1, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2))
    print("   norm_act_pad -> bf16x2 pad 1: %.1f us" % (t * 1e3))

if __name__ == "__main__":
    import subprocess
    if len(sys.argv) > 1 and sys.argv[1] == "--one":        # the trunk shape only (ncu target)
        run(1, 192, 192, 256)
    elif len(sys.argv) > 1 and sys.argv[1] == "--worker":
        run(1, 192, 192, 256)
        run(1, 128, 128, 256)
        run(1, 768, 768, 64)
        run(64, 17, 17, 64)
    else:
        for unroll in ("4", "2", "1"):
            for per_sm in ("4", "6", "8", "12"):
                print("SKIT_REDUCE_UNROLL=%s SKIT_REDUCE_BLOCKS_PER_SM=%s" % (unroll, per_sm), flush=True)
                subprocess.run([sys.executable, __file__, "--one"], env=dict(os.environ, SKIT_REDUCE_BLOCKS_PER_SM=per_sm, SKIT_REDUCE_UNROLL=unroll))
