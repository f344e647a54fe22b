"""Per-CTA phase timeline of the halo-tile tcgen05 conv (clock64 stamps via skit_debug_set_buffer)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200  # noqa
from vts_b200 import ops, _lib as L

def run(ci, co, k, s, n=1, stats=True):
    x = torch.randn(n, s, s, ci, device="cuda")
    w = torch.randn(co, ci, k, k, device="cuda") / math.sqrt(ci * k * k)
    _, op = ops.norm_act_pad(x, pad=k // 2, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2)
    pk = ops.PackedWeights(w, 0, want_f32=False, want_bf16=True)
    y = torch.empty(n, s, s, co, device="cuda")
    sm = ops.NORM_INSTANCE if stats else ops.NORM_NONE
    for _ in range(3):
        ops.conv2d_fwd(op, pk, 1, 0, s, s, stats_mode=sm, impl=ops.IMPL_TC, out=y)
    buf = torch.zeros(4096, 8, dtype=torch.int64, device="cuda")
    L.call("skit_debug_set_buffer", L.ptr(buf))
    ops.conv2d_fwd(op, pk, 1, 0, s, s, stats_mode=sm, impl=ops.IMPL_TC, out=y)
    torch.cuda.synchronize()
    L.call("skit_debug_set_buffer", None)
    b = buf.cpu()
    b = b[b[:, 0] != 0]
    d = (b[:, 1:7] - b[:, 0:1]).double()
    names = ["setup", "firstA", "mma_issued", "acc_ready", "stores_done", "end"]
    print("conv ci=%d co=%d k=%d %dx%d n=%d stats=%s: %d CTAs; mean cycles since CTA start:" % (ci, co, k, s, s, n, stats, b.shape[0]))
    print("   " + "  ".join("%s=%.0f" % (nm, d[:, i].mean().item()) for i, nm in enumerate(names)))

run(256, 256, 3, 128)
run(256, 256, 3, 128, stats=False)
run(256, 256, 3, 192)
run(64, 128, 3, 512)
