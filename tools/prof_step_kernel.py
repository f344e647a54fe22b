"""Run a few eager train steps of the headline configuration — the target of single-kernel `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/<name> python tools/prof_step_kernel.py [size]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa: E402
from oracle import skit_oracle as O  # noqa: E402  (synthetic batch factory only)

size = int(sys.argv[1]) if len(sys.argv) > 1 else 768
torch.manual_seed(0)
m = vts_b200.SinSKITGModel(vts_b200.default_options(lambda_NCE=1.0, cuda_graph=False))
m.set_input(O.synthetic_batch(size, NT=64, seed=0))
for _ in range(2):
    m.optimize_parameters(1)
torch.cuda.synchronize()
print("ok")
