"""End-to-end example on synthetic data: the reference's training loop body (train.py:37-83) with the device dataset and the B200
model — what a user of the reference runs after switching over.

    python examples/train_synthetic.py [--arch A|B] [--crop 1536] [--data-len 8] [--epochs 2]

Builds a seeded synthetic material in the reference's on-disk format (1800 x 1800 padded sketch / image / mask, 60 + 20 touch patches),
constructs `vts_b200.SingleSkitDataset` (items stay in HBM) and `vts_b200.SinSKITGModel` with the reference's option values
(arch A = its default `unet256_custom` ngf 10 / ndf 8; arch B = `resnet_9blocks` ngf 64 / ndf 64), and trains for a few epochs.
Prints the dataset build time, the steady-state step time and the last losses.  Needs a B200 and the built library.
"""
import argparse
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vts_b200  # noqa: E402
from tools.bench_data import big_dataset, options  # noqa: E402  (synthetic data factory)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="A", choices=["A", "B"])
    ap.add_argument("--crop", type=int, default=1536, help=">= 1280: the crop must cover the synthetic material's 1280 x 960 centre region (dataset_util.py:168)")
    ap.add_argument("--data-len", type=int, default=8)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--root", default="/tmp/vts_bench_data/singleskit_syn_padded_1800_x1")
    a = ap.parse_args()
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    root = big_dataset(a.root)
    dopt = options(root, a.data_len)
    dopt.crop_size = a.crop
    t0 = time.time()
    ds = vts_b200.SingleSkitDataset(dopt)
    torch.cuda.synchronize()
    print("dataset: %d augmentations built in %.2f s" % (len(ds), time.time() - t0))
    loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=True, num_workers=0, drop_last=True)
    arch = dict(netG="unet256_custom", ngf=10, ndf=8) if a.arch == "A" else dict(netG="resnet_9blocks", ngf=64, ndf=64)
    opt = vts_b200.default_options(crop_size=a.crop, **arch)
    model = vts_b200.SinSKITGModel(opt)
    model.setup(opt)
    steps, t_start = 0, None
    for epoch in range(1, a.epochs + 1):
        for data in loader:                          # train.py:56-63
            model.set_input(data)
            model.optimize_parameters(epoch)
            steps += 1
            if steps == 4:                           # past the eager warm-up steps and the graph capture
                torch.cuda.synchronize(); t_start = time.time(); s0 = steps
        model.update_learning_rate()                 # train.py:205
    torch.cuda.synchronize()
    losses = model.get_current_losses()
    assert all(np.isfinite(float(v)) for v in losses.values()), losses
    if t_start is not None and steps > s0:
        print("train: %d steps, %.2f ms/step (arch %s, %dx%d, host loop + device dataset)" % (steps, 1e3 * (time.time() - t_start) / (steps - s0), a.arch, a.crop, a.crop))
    print("losses:", {k: round(float(v), 4) for k, v in losses.items()})
    print("EXAMPLE OK")


if __name__ == "__main__":
    main()
