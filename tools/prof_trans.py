"""Run the pixels-on-N tcgen05 conv (conv_tc_halo_t_kernel) at the 768 x 768 step's largest shape — 64 -> 128, k3, at full resolution — a
few times: the target of an `ncu --set full` capture.
    ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo_t -s 4 -c 1 -o gpurun_out/prof python tools/prof_trans.py [S] [fwd|dgrad]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa: E402,F401
from vts_b200 import ops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 768
which = sys.argv[2] if len(sys.argv) > 2 else "fwd"
x = torch.randn(1, S, S, 64, device="cuda")
w = torch.randn(128, 64, 3, 3, device="cuda") / math.sqrt(64 * 9)
_, op = ops.norm_act_pad(x, pad=1, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2)
g = torch.randn(1, S, S, 128, device="cuda")
_, dop = ops.norm_act_pad(g, pad=2, pad_mode=ops.PAD_ZERO, fmt=ops.FMT_BF16X2)
pk = ops.PackedWeights(w, 0, want_f32=False, want_bf16=True)
pk1 = ops.PackedWeights(w, 1, want_f32=False, want_bf16=True)
for _ in range(6):
    if which == "fwd":
        ops.conv2d_fwd(op, pk, 1, 0, S, S, stats_mode=ops.NORM_INSTANCE, impl=ops.IMPL_TC)
    else:
        ops.conv2d_dgrad_s1(dop, pk1)
torch.cuda.synchronize()
print("ok")
