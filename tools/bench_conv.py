"""Micro-benchmark of the conv kernels at the generator's real layer shapes (CUDA events)."""
import math
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa: E402
from vts_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=3):
    """Device time per call: the calls are captured into one CUDA graph so that host-side launch cost
    (ctypes, tensor-map encoding) does not bound the measurement of short kernels."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    shapes = [(256, 256, 3, 128, 1), (256, 256, 3, 192, 1), (256, 256, 3, 256, 1), (64, 128, 3, 512, 1),
              (128, 256, 3, 256, 1), (128, 64, 3, 512, 1), (256, 256, 3, 128, 4)]
    for ci, co, k, s, n in shapes:
        x = torch.randn(n, s, s, ci, device="cuda")
        w = torch.randn(co, ci, k, k, device="cuda") / math.sqrt(ci * k * k)
        _, op = ops.norm_act_pad(x, pad=k // 2, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2)
        pk = ops.PackedWeights(w, 0, want_f32=True, want_bf16=True)
        y = torch.empty(n, s, s, co, device="cuda")
        fl = 2.0 * k * k * ci * co * s * s * n
        for impl, name in ((ops.IMPL_TC, "tcgen05"), (ops.IMPL_SIMT, "simt")):
            if impl == ops.IMPL_SIMT and s > 256:
                continue
            ms = timeit(lambda: ops.conv2d_fwd(op, pk, 1, 0, s, s, stats_mode=ops.NORM_INSTANCE, impl=impl, out=y), iters=10 if impl == ops.IMPL_TC else 3)
            print("conv %s ci=%d co=%d k=%d %dx%d n=%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (name, ci, co, k, s, s, n, ms, fl / ms / 1e9), flush=True)
        ms = timeit(lambda: ops.norm_act_pad(x, pad=1, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2))
        print("  prep(pad+split) %.3f ms  %.0f GB/s" % (ms, x.numel() * 8 / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
