"""What bounds the dominant kernel?  Times conv_tc_halo_kernel<256> (ResnetBlock conv, forward three-term and input-gradient
two-term) as is, with no MMAs issued (SKIT_DBG_MODE=1: the TMA load pipelines alone), with no filter TMA (2: MMA + activation
loads) and with no activation TMA (3).  Debug modes compute garbage; each runs in a fresh process.   python tools/bench_limiter.py [size]"""
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
    s = size // 4
    nbuf = 4
    w = torch.randn(256, 256, 3, 3, device="cuda") / math.sqrt(2304)
    pk, pk1 = ops.PackedWeights(w, 0, want_f32=False, want_bf16=True), ops.PackedWeights(w, 1, want_f32=False, want_bf16=True)
    xs = [ops.norm_act_pad(torch.randn(1, s, s, 256, device="cuda"), pad=1, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2)[1] for _ in range(nbuf)]
    ds = [ops.norm_act_pad(torch.randn(1, s, s, 256, device="cuda"), pad=2, pad_mode=ops.PAD_ZERO, fmt=ops.FMT_BF16X2)[1] for _ in range(nbuf)]
    ys = [torch.empty(1, s, s, 256, device="cuda") for _ in range(nbuf)]
    i = [0]

    def fwd():
        i[0] += 1
        ops.conv2d_fwd(xs[i[0] % nbuf], pk, 1, 0, s, s, stats_mode=ops.NORM_INSTANCE, impl=ops.IMPL_TC, out=ys[i[0] % nbuf])

    def dgrad():
        i[0] += 1
        ops.conv2d_dgrad_s1(ds[i[0] % nbuf], pk1)

    print("   fwd (3-term) %.1f us | dgrad (2-term) %.1f us" % (1e3 * timeit(fwd, iters=12), 1e3 * timeit(dgrad, iters=12)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]))
        sys.exit(0)
    size = sys.argv[1] if len(sys.argv) > 1 else "768"
    for mode, what in (("0", "as is"), ("1", "no MMAs (loads alone)"), ("2", "no filter TMA"), ("3", "no activation TMA")):
        print("SKIT_DBG_MODE=%s  %s" % (mode, what), flush=True)
        subprocess.run([sys.executable, __file__, "--worker", size], env=dict(os.environ, SKIT_DBG_MODE=mode, SKIT_TC_TRANS="0"))
