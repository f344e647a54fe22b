"""Micro-benchmark of the <= 128-output-channel conv layers at full resolution: the pixel-major kernels (SKIT_TC_TRANS=0) against
the pixels-on-N kernel in its tile shapes (SKIT_TRANS_TY x SKIT_TRANS_NA), forward (three-term) and input gradient (two-term).
Each configuration runs in a fresh process because the library caches its environment switches.
    python tools/bench_trans.py [size]"""
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(size):
    import torch
    sys.path.insert(0, ROOT)
    import vts_b200  # noqa: F401
    from vts_b200 import ops
    from tools.bench_conv import timeit
    for ci, co, k, s in ((256, 256, 3, size // 4), (64, 128, 3, size), (128, 64, 3, size), (256, 128, 3, size // 2), (128, 256, 3, size // 2)):
        x = torch.randn(1, s, s, ci, device="cuda")
        w = torch.randn(co, ci, 3 if k == 7 else k, 3 if k == 7 else k, device="cuda") / math.sqrt(ci * 9)
        kk = 3 if k == 7 else k
        _, op = ops.norm_act_pad(x, pad=kk // 2, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2)
        pk = ops.PackedWeights(w, 0, want_f32=False, want_bf16=True)
        y = torch.empty(1, s, s, co, device="cuda")
        fl = 2.0 * kk * kk * ci * co * s * s
        ms = timeit(lambda: ops.conv2d_fwd(op, pk, 1, 0, s, s, stats_mode=ops.NORM_INSTANCE, impl=ops.IMPL_TC, out=y), iters=8)
        # input gradient of the same layer: K = co, N = ci
        _, dop = ops.norm_act_pad(torch.randn(1, s, s, co, device="cuda"), pad=kk - 1, pad_mode=ops.PAD_ZERO, fmt=ops.FMT_BF16X2)
        pk1 = ops.PackedWeights(w, 1, want_f32=False, want_bf16=True)
        ms_d = timeit(lambda: ops.conv2d_dgrad_s1(dop, pk1), iters=8)
        print("  %3d->%3d k%d @%d: fwd %.3f ms %.0f TFLOP/s | dgrad %.3f ms %.0f TFLOP/s (algorithmic)"
              % (ci, co, kk, s, ms, fl / ms / 1e9, ms_d, fl / ms_d / 1e9), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]))
        sys.exit(0)
    size = sys.argv[1] if len(sys.argv) > 1 else "768"
    for env in ({"SKIT_TC_TRANS": "0"}, {}, {"SKIT_TRANS_MIN_PM_TILES": "0"}, {"SKIT_TRANS_MIN_PM_TILES": "0", "SKIT_DGRAD_SPLIT": "0"},
                {"SKIT_TRANS_MIN_PM_TILES": "0", "SKIT_TRANS_TY": "24", "SKIT_TRANS_NA": "2"}):
        print(env or "default", flush=True)
        subprocess.run([sys.executable, __file__, "--worker", size], env=dict(os.environ, **env))
