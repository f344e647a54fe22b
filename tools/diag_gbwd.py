"""Diagnostic: per-parameter gradient error of the CUDA generator backward vs the fp32 and fp64 CPU oracle."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vts_b200
from oracle import skit_oracle as O

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

def rand_input(seed, *shape):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * 2 - 1

def main(hw=(32, 32), n=1, nb=9, gain=1.0):
    torch.manual_seed(3)
    G = vts_b200.define_G(9, 5, 64, "resnet_%dblocks" % nb, "instance", False, "xavier", gain, False, False, [], argparse.Namespace())
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    G = G.cuda(); G.ensure_flat(); G.refresh_packs()
    x = rand_input(5, n, 9, *hw); M = (rand_input(6, n, 1, *hw) > -0.8).float()
    RI, RT = rand_input(7, n, 3, *hw), rand_input(8, n, 2, *hw)
    grads = {}
    for dt in (torch.float32, torch.float64):
        ps = {k: v.clone().to(dt).requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "filt" not in k}
        run = {k: v.to(dt) if v.dtype.is_floating_point else v for k, v in sd.items()}
        run.update(ps)
        O._BLUR3, O._BLUR4 = O._BLUR3.to(dt), O._BLUR4.to(dt)
        out = O.resnet_g_forward(run, x.to(dt), n_blocks=nb)
        fI, fT = out[:, :3] * M.to(dt), out[:, 3:] * M.to(dt)
        ((fI * RI.to(dt)).sum() + (fT * RT.to(dt)).sum()).backward()
        grads[dt] = ({k: p.grad for k, p in ps.items()}, out.detach())
    (kI, kT, kN), ctx, _ = G.fwd([x.cuda()], mask=M.cuda())
    G.zero_grad(); G.bwd(ctx, RI.cuda(), RT.cuda()); torch.cuda.synchronize()
    o64 = grads[torch.float64][1]
    print("fwd: cuda vs f64 %.2e | cpu f32 vs f64 %.2e" % (rel(torch.cat([kI, kT], 1), o64 * M), rel(grads[torch.float32][1], o64)))
    for k, p in G.named_parameters():
        if k.endswith("bias"):
            continue
        g32, g64 = grads[torch.float32][0][k], grads[torch.float64][0][k]
        print("%-34s cuda-vs-f64 %.2e  cpu32-vs-f64 %.2e  cuda-vs-cpu32 %.2e" % (k, rel(p.grad, g64), rel(g32, g64), rel(p.grad, g32)))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        main((64, 64), 1, int(sys.argv[1]), 0.02)
    else:
        main()
        main((64, 48), 2, 4)
        main((64, 64), 1, 9, 0.02)
