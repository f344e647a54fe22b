"""Run the tcgen05 ResnetBlock conv (256->256 k3) a few times — the target of `ncu --set full` captures.
    ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 6 -c 2 -o gpurun_out/prof python tools/prof_conv.py [S] [which]
which: fwd | dgrad | wgrad"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa: E402,F401
from vts_b200 import ops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
which = sys.argv[2] if len(sys.argv) > 2 else "fwd"
s = S // 4
x = torch.randn(1, s, s, 256, device="cuda")
w = torch.randn(256, 256, 3, 3, device="cuda") / math.sqrt(2304)
_, op = ops.norm_act_pad(x, pad=1, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2)
_, dop = ops.norm_act_pad(x, pad=2, pad_mode=ops.PAD_ZERO, fmt=ops.FMT_BF16X2)
_, dop0 = ops.norm_act_pad(x, pad=0, pad_mode=ops.PAD_ZERO, fmt=ops.FMT_BF16X2)
pk = ops.PackedWeights(w, 0, want_f32=False, want_bf16=True)
pk1 = ops.PackedWeights(w, 1, want_f32=False, want_bf16=True)
dw = torch.zeros_like(w)
for _ in range(8):
    if which == "fwd":
        ops.conv2d_fwd(op, pk, 1, 0, s, s, stats_mode=ops.NORM_INSTANCE, impl=ops.IMPL_TC)
    elif which == "dgrad":
        ops.conv2d_dgrad_s1(dop, pk1)     # interior + border-strip regions in one grid
    else:
        ops.conv2d_wgrad(op, 0, dop0, 0, 3, 1, s, s, dw, None, impl=ops.IMPL_TC)
torch.cuda.synchronize()
print("ok")
