"""Parity of the CUDA path against the CPU oracle AT BASELINE.json's own configurations (VERDICT r1 'Next round' item 1):
  (a) arch-B generator forward at 512x512 and 768x768 vs O.resnet_g_forward / O.model_forward — fake_I / fake_T / fake_N 1e-3 relative;
  (b) one full train step at 512x512 (configs[1]: NT 64, NF 32) and at 768x768 with PatchNCE (configs[2]) vs O.train_step — every
      loss 1e-3, every gradient tensor by the robust metric of test_train_step_gpu.py (worst value per net printed);
  (c) the conv kernels at the shapes that dispatch the persistent halo kernel (>= 296 units: 64 -> 128 at 256x256, 128 -> 64),
      the two-wave grids (256 -> 256 at 192x192 = 288 tiles) and the multi-region input gradient (130x130 trunk gradient: interior
      + four border strips) directly against torch's fp32 convolution.
The oracle needs seconds (forward) to tens of seconds (768x768 step) of host time per case."""
import copy
import math
import time

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GATE = 1e-3
POST_UPDATE_GATE = 1e-2
GRAD_REL, GRAD_COS = 3e-2, 0.9995


def rel(a, b):
    b = b.detach().cpu() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b))
    a, b = a.detach().double().cpu(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cos(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def check_grads(net, ref_grads, tag):
    worst_r, worst_c, worst_k = 0.0, 1.0, None
    for k, p in net.named_parameters():
        if k not in ref_grads:
            continue
        g_ref = ref_grads[k]
        wk = k.replace(".bias", ".weight")
        if k.endswith(".bias") and wk in ref_grads and g_ref.norm() < 1e-3 * ref_grads[wk].norm():
            continue  # conv bias feeding a norm layer: mathematically zero, the reference holds rounding noise
        r, c = rel(p.grad, g_ref), cos(p.grad, g_ref)
        if r > worst_r:
            worst_r, worst_k = r, k
        worst_c = min(worst_c, c)
        assert r < GRAD_REL and c > GRAD_COS, (tag, k, r, c)
    print("%s: worst gradient rel err %.3e (%s), worst cosine %.6f" % (tag, worst_r, worst_k, worst_c))


@pytest.mark.parametrize("S,scale", [(512, 20.0), (768, 1.0)])
def test_generator_forward_vs_oracle_at_baseline_sizes(S, scale):
    """SinSKITGModel.test() at the BASELINE sizes: xavier(0.02) weights as initialised (768) and scaled x20 so that the tanh head
    leaves its linear range (512).  Non-trivial background mask."""
    import vts_b200
    from oracle import skit_oracle as O
    torch.manual_seed(S)
    m = vts_b200.SinSKITGModel(vts_b200.default_options(isTrain=False))
    m.netG.ensure_flat()
    with torch.no_grad():
        m.netG.flat_param.mul_(scale)
    m.netG.refresh_packs()
    sd = {k: v.detach().cpu().clone() for k, v in m.netG.state_dict().items()}
    b = O.synthetic_batch(S, NT=4, seed=3, ellipse_mask=True)
    m.set_input({"S": b["S"], "M": b["M"]}, phase="test")
    fI, fT, fN = m.test()
    torch.cuda.synchronize()
    t0 = time.time()
    ref = O.model_forward(O.StepConfig(), sd, b["S"] * b["M"], O.spe_grid(S, S, 4, 1), b["M"])
    errs = {k: rel(v, ref[k]) for k, v in (("fake_I", fI), ("fake_T", fT), ("fake_N", fN))}
    print("forward %dx%d vs oracle (%.1f s of CPU): %s" % (S, S, time.time() - t0, errs))
    assert max(errs.values()) < GATE, errs
    fI2, _, _ = m.test()    # the captured-graph replay computes the same thing
    fI3, _, _ = m.test()
    torch.cuda.synchronize()
    assert rel(fI3, ref["fake_I"]) < GATE


@pytest.mark.parametrize("S,nce", [(512, False), (768, True)])
def test_train_step_vs_oracle_at_baseline_configs(S, nce):
    """BASELINE.json configs[1] (512x512, PatchNCE off) and configs[2] (768x768, PatchNCE + GAN): NT = 64 touch patches, NF = 32
    random fake patches, arch B.  Losses, generator outputs and every gradient bucket entry against the oracle's step."""
    import vts_b200
    from oracle import skit_oracle as O
    NT, NF, P = 64, 32, 256
    torch.manual_seed(7)
    opt = vts_b200.default_options(lambda_NCE=1.0 if nce else 0.0, num_patches=P, cuda_graph=False)
    m = vts_b200.SinSKITGModel(opt)
    sds = [{k: v.detach().cpu().clone() for k, v in net.state_dict().items()} for net in (m.netG, m.netD, m.netD2)]
    batch = O.synthetic_batch(S, NT=NT, seed=1, ellipse_mask=True)
    rs = np.random.RandomState(5)
    rand = dict(real_b=[0.3], real_s=[0.8], fake_b=[0.6], fake_s=[0.2],
                fake_ox=rs.randint(0, S - 32, NF).astype(np.int32), fake_oy=rs.randint(0, S - 32, NF).astype(np.int32))
    if nce:
        sizes = [m.netG.feature_hw(l, S, S) for l in m.nce_layers]
        rand["nce_ids"] = [rs.permutation(h * w)[:min(P, h * w)] for h, w in sizes]
    m.set_input(batch)
    m.optimize_parameters(1, rand=rand)
    torch.cuda.synchronize()
    losses = m.current_losses()
    cfg = O.StepConfig(netG="resnet_9blocks", batch_size_G2=NT, add_fake_T_sample_size=NF, lambda_NCE=1.0 if nce else 0.0, num_patches=P)
    sdG, sdD, sdD2 = [copy.deepcopy(s) for s in sds]
    t0 = time.time()
    res = O.train_step(cfg, sdG, sdD, sdD2, {}, O.step_inputs_from_batch(batch), rand, step=1)
    print("oracle train step %dx%d%s: %.1f s of CPU" % (S, S, " + PatchNCE" if nce else "", time.time() - t0))
    assert ("NCE" in res["losses"]) == nce
    for k, v in res["losses"].items():
        # G_GAN / G2_GAN are evaluated through discriminators that were just updated by Adam with beta1 = 0 at step 1, i.e. by
        # -lr * sign(g) on every weight.  Two fp32 implementations of the forward differ at the 1e-5 level, which flips ~1e-5 of
        # the ReLU / LeakyReLU masks and moves weight gradients by ~sqrt(1e-5) = 3e-3 (the reason gradients are gated at 3e-2),
        # so a fraction of a percent of the weight-gradient SIGNS differ and these two values move in their third digit —
        # tools/step_conditioning.py reproduces it on the CPU oracle alone (weights perturbed by 1e-4: G_GAN moves by 1.2e-3 at
        # 256x256).  Measured here: 1.1e-3 at 512x512, 5.4e-3 at 768x768.  Everything computed before an update keeps 1e-3.
        gate = POST_UPDATE_GATE if k in ("G_GAN", "G2_GAN") else GATE
        assert abs(losses[k] - v) <= gate * max(1.0, abs(v)), (k, losses[k], v)
    print("losses: max rel deviation %.2e" % max(abs(losses[k] - v) / max(1.0, abs(v)) for k, v in res["losses"].items()))
    errs = {nm: rel(getattr(m, nm), res[nm]) for nm in ("fake_I", "fake_T", "fake_N", "aug_fake_I")}
    print("outputs vs oracle:", errs)
    assert max(errs.values()) < GATE, errs
    for tag, net, grads in (("G", m.netG, res["grads_G"]), ("D", m.netD, res["grads_D"]), ("D2", m.netD2, res["grads_D2"])):
        check_grads(net, grads, "%s @%d" % (tag, S))


def _conv_case(V, ci, co, k, hw, pad, mode, seed):
    ops = V.ops
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, ci, *hw, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) / math.sqrt(ci * k * k)
    b = torch.randn(co, generator=g)
    xp = F.pad(x, (pad,) * 4, mode={0: "constant", 1: "reflect"}[mode]) if pad else x
    _, op = ops.norm_act_pad(x.permute(0, 2, 3, 1).contiguous().cuda(), pad=pad, pad_mode=mode, fmt=ops.FMT_BF16X2)
    return x, xp, w, b, op


@pytest.mark.parametrize("ci,co,hw,what", [
    (64, 128, (256, 256), "persistent N=128, 512 tiles"),
    (128, 64, (256, 256), "persistent N=64, 512 tiles"),
    (128, 256, (128, 128), "N=256 single wave, 128 tiles"),
    (256, 256, (192, 192), "N=256 two waves, 288 tiles (the 768x768 trunk)"),
    (256, 256, (136, 120), "N=256 ragged tiles"),
])
def test_conv_fwd_tcgen05_at_full_size_shapes(ci, co, hw, what):
    import vts_b200 as V
    ops = V.ops
    x, xp, w, b, op = _conv_case(V, ci, co, 3, hw, 1, 1, ci + co + hw[0])
    ref = F.conv2d(xp, w, b)
    pk = ops.PackedWeights(w.cuda(), 0, want_f32=False, want_bf16=True)
    y, st = ops.conv2d_fwd(op, pk, 1, 0, hw[0], hw[1], bias=b.cuda(), stats_mode=ops.NORM_INSTANCE, impl=ops.IMPL_TC)
    torch.cuda.synchronize()
    r = rel(y.permute(0, 3, 1, 2), ref)
    print("%s: rel err %.2e" % (what, r))
    assert r < 5e-5, what
    s = st.cpu()[0]
    np.testing.assert_allclose(s[:, 0], ref.double().sum((0, 2, 3)), rtol=1e-4, atol=5e-2)
    np.testing.assert_allclose(s[:, 1], (ref.double() ** 2).sum((0, 2, 3)), rtol=1e-4, atol=5e-2)


@pytest.mark.parametrize("terms,tol", [(3, 5e-5), (2, 4e-3)])
@pytest.mark.parametrize("c,hw", [(256, (128, 128)), (256, (192, 192)), (128, (256, 256))])
def test_dgrad_s1_multi_region_at_full_size_shapes(c, hw, terms, tol):
    """Stride-1 input gradient (interior + four border strips in one launch) of a 3x3 conv at the trunk / up-conv sizes vs
    torch's conv_transpose2d (= the autograd input gradient of the VALID conv over the haloed operand)."""
    import vts_b200 as V
    ops = V.ops
    g = torch.Generator().manual_seed(c + hw[0])
    dy = torch.randn(1, c, *hw, generator=g)
    w = torch.randn(c, c, 3, 3, generator=g) / math.sqrt(c * 9)
    ref = F.conv_transpose2d(dy, w)                       # [1, c, h + 2, w + 2]
    _, d_op = ops.norm_act_pad(dy.permute(0, 2, 3, 1).contiguous().cuda(), pad=2, pad_mode=ops.PAD_ZERO, fmt=ops.FMT_BF16X2)
    pk = ops.PackedWeights(w.cuda(), 1, want_f32=False, want_bf16=True)
    ops.set_backward_terms(terms)       # 3: the forward's full product; 2: the default backward precision (W_hi only)
    try:
        dx = ops.conv2d_dgrad_s1(d_op, pk)
        torch.cuda.synchronize()
    finally:
        ops.set_backward_terms(2)
    assert tuple(dx.shape) == (1, hw[0] + 2, hw[1] + 2, c)
    r = rel(dx.permute(0, 3, 1, 2), ref)
    print("dgrad_s1 %d ch @%dx%d, %d-term product: rel err %.2e" % (c, hw[0], hw[1], terms, r))
    assert r < tol


@pytest.mark.parametrize("ci,co,k,hw,n", [
    (64, 128, 3, (250, 203), 1),      # ragged tiles in both directions, N = 192 pixels per instruction (three-term forward)
    (128, 64, 3, (100, 100), 3),      # 64 output channels (rows 64..127 of the M tile are TMA zero fill), several images
    (192, 96, 3, (120, 136), 1),      # channel counts that are not powers of two: 3 K chunks, 96 of 128 M rows
    (64, 64, 4, (150, 160), 2),       # even filter (PatchGAN-style 4x4, stride 1)
])
def test_conv_fwd_pixels_on_n_kernel(ci, co, k, hw, n):
    """conv_tc_halo_t_kernel (channels on M, 192 / 256 pixels on N: layers with <= 128 output channels on maps that fill the
    SMs) against torch's fp32 convolution: values, bias, and the per-image InstanceNorm statistics of its epilogue."""
    import vts_b200 as V
    ops = V.ops
    g = torch.Generator().manual_seed(ci + co + k)
    pad = k // 2
    x = torch.randn(n, ci, *hw, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) / math.sqrt(ci * k * k)
    b = torch.randn(co, generator=g)
    ref = F.conv2d(F.pad(x, (pad,) * 4, mode="reflect"), w, b)
    _, op = ops.norm_act_pad(x.permute(0, 2, 3, 1).contiguous().cuda(), pad=pad, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2)
    pk = ops.PackedWeights(w.cuda(), 0, want_f32=False, want_bf16=True)
    ho, wo = ref.shape[2:]
    y, st = ops.conv2d_fwd(op, pk, 1, 0, ho, wo, bias=b.cuda(), stats_mode=ops.NORM_INSTANCE, impl=ops.IMPL_TC)
    torch.cuda.synchronize()
    r = rel(y.permute(0, 3, 1, 2), ref)
    print("pixels-on-N conv %d->%d k%d @%dx%d n=%d: rel err %.2e" % (ci, co, k, hw[0], hw[1], n, r))
    assert r < 5e-5
    s = st.cpu()
    np.testing.assert_allclose(s[..., 0], ref.double().sum((2, 3)), rtol=1e-4, atol=5e-2)
    np.testing.assert_allclose(s[..., 1], (ref.double() ** 2).sum((2, 3)), rtol=1e-4, atol=5e-2)


@pytest.mark.parametrize("terms,tol", [(3, 5e-5), (2, 4e-3)])
def test_dgrad_s2_parity_subconvs_on_large_map(terms, tol):
    """Stride-2 input gradient (PatchGAN 64 -> 128, k4 s2 at a 400x400 input: the four parity sub-convolutions write interleaved
    quarters of dx through the pixels-on-N kernel's strided output placement) vs torch's conv_transpose2d."""
    import vts_b200 as V
    ops = V.ops
    g = torch.Generator().manual_seed(11)
    ci, co, k, H = 64, 128, 4, 400
    hp = H + 4                                   # zero halo of 2, as the discriminators pad (networks.py:1703)
    ho = (hp - k) // 2 + 1
    dy = torch.randn(1, co, ho, ho, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) / math.sqrt(ci * k * k)
    ref = F.conv_transpose2d(dy, w, stride=2)     # gradient w.r.t. the padded input: (ho - 1) * 2 + 4 = hp
    assert ref.shape[-1] == hp
    _, d_op = ops.norm_act_pad(dy.permute(0, 2, 3, 1).contiguous().cuda(), pad=1, pad_mode=ops.PAD_ZERO, fmt=ops.FMT_BF16X2)
    pk = ops.PackedWeights(w.cuda(), 3, want_f32=False, want_bf16=True)
    ops.set_backward_terms(terms)
    try:
        dx = ops.conv2d_dgrad_s2(d_op, 1, pk, k, ho, ho, hp, hp)
        torch.cuda.synchronize()
    finally:
        ops.set_backward_terms(2)
    r = rel(dx.permute(0, 3, 1, 2), ref)
    print("dgrad_s2 %d-term: rel err %.2e" % (terms, r))
    assert r < tol
