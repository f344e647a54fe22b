"""Host-side phases of the training loop with the device dataset (fetch + collate, set_input, train step), synchronised per phase:
    python tools/prof_e2e_loop.py      # 1536 x 1536, arch B, 4 augmentations x 4 epochs"""
import os, sys, time, random
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import vts_b200
from tools.bench_data import big_dataset, options
random.seed(0); np.random.seed(0); torch.manual_seed(0)
root = big_dataset("/tmp/vts_bench_data/singleskit_syn_padded_1800_x1")
dopt = options(root, 4); dopt.crop_size = 1536
ds = vts_b200.SingleSkitDataset(dopt)
loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=0, drop_last=True)
opt = vts_b200.default_options(crop_size=1536, netG="resnet_9blocks", ngf=64, ndf=64)
m = vts_b200.SinSKITGModel(opt); m.setup(opt)
def sync(): torch.cuda.synchronize(); return time.time()
for ep in range(4):
    it = iter(loader)
    for k in range(len(loader)):
        t0 = sync(); data = next(it); t1 = sync(); m.set_input(data); t2 = sync(); m.optimize_parameters(ep + 1); t3 = sync()
        print("ep %d item %d: fetch+collate %.1f ms, set_input %.1f ms, step %.1f ms" % (ep, k, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)))
