"""Graph-timed forward of the StyleGAN2 generator (ngf 64) next to the resnet_9blocks generator, per image size."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa: E402
from tools.bench_conv import timeit  # noqa: E402


def main():
    for S in (512, 1024):
        opt = argparse.Namespace(load_size=S, crop_size=S, stylegan2_G_num_downsampling=1, netG="stylegan2")
        G = vts_b200.networks.define_G(9, 5, 64, "stylegan2", "instance", False, "xavier", 0.02, False, False, [0], opt)
        x = torch.rand(1, 9, S, S, device="cuda") * 2 - 1
        nz = [torch.randn(1, 1, S, S, device="cuda")]
        G(x, noises=nz)
        ms = timeit(lambda: G(x, noises=nz), iters=5)
        # algorithmic flops (reference formulation: 3x3 s2 conv after the blur, conv_transpose 3x3)
        c0, c1 = 64, 128
        fl = 2 * S * S * 9 * c0 + 2 * S * S * 9 * c0 * c0 + 2 * (S // 2) ** 2 * (9 * c0 * c1 + c0 * c1)
        fl += 12 * 2 * (S // 2) ** 2 * 9 * c1 * c1 + 2 * (S // 2) ** 2 * 9 * c1 * c0 + 2 * S * S * c0 * 3
        print("stylegan2 ngf64 %dx%d forward: %.3f ms  (%.1f images/s, %.1f algorithmic TFLOP/s)" % (S, S, ms, 1e3 / ms, fl / ms / 1e9), flush=True)
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            G(x, noises=nz)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]
        for e in rows:
            print("   %8.1f us x%-3d %s" % (e.device_time_total, e.count, e.key[:90]))


if __name__ == "__main__":
    main()
