"""GPU parity of the full train step (SinSKITGModel.optimize_parameters on the B200 path) against
(1) the golden fixture produced by the REAL reference's optimize_parameters (tests/golden/step_resnet.npz,
written by oracle/make_golden.py) and (2) the CPU oracle at the tensor-core configuration (ngf = ndf = 64).
Gates: outputs and losses 1e-3 relative (north star); gradients by the robust metric explained in
test_networks_gpu.py (mask flips), post-Adam weights where the gradient is not sign-ambiguous."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GATE = 1e-3
GRAD_REL, GRAD_COS = 3e-2, 0.9995


def rel(a, b):
    b = b.detach().cpu() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b))
    a, b = a.detach().double().cpu(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cos(a, b):
    b = b.detach().cpu() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b))
    a, b = a.detach().double().cpu().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def sd_from(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(prefix)}


def check_grads(net, ref_grads, tag):
    worst = 0.0
    for k, p in net.named_parameters():
        if k not in ref_grads:
            continue
        g_ref = torch.as_tensor(np.asarray(ref_grads[k]))
        wk = k.replace(".bias", ".weight")
        if k.endswith(".bias") and wk in ref_grads and g_ref.norm() < 1e-3 * torch.as_tensor(np.asarray(ref_grads[wk])).norm():
            continue  # conv bias feeding a norm layer: mathematically zero, the reference holds rounding noise
        r, c = rel(p.grad, g_ref), cos(p.grad, g_ref)
        worst = max(worst, r)
        assert r < GRAD_REL and c > GRAD_COS, (tag, k, r, c)
    return worst


@pytest.mark.parametrize("tag,netG,wkey", [("resnet", "resnet_9blocks", "model.1.weight"), ("unet", "unet256_custom", "down0.model.0.weight")])
def test_train_step_matches_reference_golden(golden_dir, tag, netG, wkey):
    import vts_b200
    from oracle import skit_oracle as O
    z = np.load(os.path.join(golden_dir, "step_%s.npz" % tag))
    S, NT, NF = [int(v) for v in z["meta"]]
    sdG, sdD, sdD2 = sd_from(z, "G_before."), sd_from(z, "D_before."), sd_from(z, "D2_before.")
    opt = vts_b200.default_options(netG=netG, ngf=sdG[wkey].shape[0], ndf=sdD["layer0.0.weight"].shape[0],
                                   batch_size_G2=NT, add_fake_T_sample_size=NF, run_full_res_D2=True)
    m = vts_b200.SinSKITGModel(opt)
    m.netG.load_state_dict(sdG)
    m.netD.load_state_dict(sdD)
    m.netD2.load_state_dict(sdD2)
    m.set_input(O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True))
    u = z["rand_u"]
    rand = dict(real_b=u[0], real_s=u[1], fake_b=u[2], fake_s=u[3], fake_ox=z["fake_ox"], fake_oy=z["fake_oy"])
    m.optimize_parameters(1, rand=rand)
    torch.cuda.synchronize()
    losses = m.current_losses()
    for k, v in losses.items():
        ref = float(z["loss_l_" + k])
        assert abs(v - ref) <= GATE * max(1.0, abs(ref)), (k, v, ref)
    for nm in ("fake_I", "fake_T", "fake_N", "aug_fake_I"):
        assert rel(getattr(m, nm)[..., ::2, ::2], z[nm]) < GATE, nm
    assert rel(m.pred_fake_T_full.permute(0, 3, 1, 2), z["pred_fake_T_full"]) < GATE
    for tag, net in (("G", m.netG), ("D", m.netD), ("D2", m.netD2)):
        grads = {k[len(tag) + 6:]: z[k] for k in z.files if k.startswith(tag + "_grad.")}
        assert grads
        w = check_grads(net, grads, tag)
        print(tag, "worst grad rel err vs reference", w)
        sd_now = net.state_dict()
        for k in sd_now:
            sk, gk = "%s_after_sub.%s" % (tag, k), "%s_grad.%s" % (tag, k)
            if sk in z.files and gk in z.files:
                g = z[gk].reshape(-1)[::7]
                ok = np.abs(g) > max(1e-6, 0.05 * float(np.sqrt(np.mean(z[gk] ** 2))))
                wk = "%s_grad.%s" % (tag, k.replace(".bias", ".weight"))
                if k.endswith(".bias") and wk in z.files and np.linalg.norm(z[gk]) < 1e-3 * np.linalg.norm(z[wk]):
                    continue
                if not ok.any():
                    continue
                mine = sd_now[k].detach().cpu().reshape(-1)[::7].numpy()[ok]
                # beta1 = 0, first step: every weight moves by ~lr*sign(g); compare where the sign is unambiguous,
                # allowing the few elements whose gradient sign flipped with a ReLU mask (see test_networks_gpu.py)
                bad = np.abs(mine - z[sk][ok]) > 1e-4 * np.abs(z[sk][ok]) + 2e-6
                assert bad.mean() < 0.01, (tag, k, bad.mean())
            ak = "%s_after.%s" % (tag, k)
            if ak in z.files and sd_now[k].dtype.is_floating_point:
                np.testing.assert_allclose(sd_now[k].detach().cpu().numpy(), z[ak], rtol=1e-3, atol=3e-4)


@pytest.mark.parametrize("gan_mode", ["nonsaturating", "hinge"])
def test_train_step_ngf64_tcgen05_vs_oracle(gan_mode):
    """Tensor-core configuration (arch B: resnet_9blocks ngf 64 + multiscale ndf 64) at S = 64, in both GAN modes the reference's
    step can run (per-sample losses: 'nonsaturating', the default, and 'hinge')."""
    import vts_b200
    from oracle import skit_oracle as O
    S, NT, NF = 64, 8, 4
    torch.manual_seed(1)
    opt = vts_b200.default_options(batch_size_G2=NT, add_fake_T_sample_size=NF, run_full_res_D2=True, gan_mode=gan_mode)
    m = vts_b200.SinSKITGModel(opt)
    sds = [{k: v.detach().cpu().clone() for k, v in net.state_dict().items()} for net in (m.netG, m.netD, m.netD2)]
    batch = O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True)
    rand = dict(real_b=[0.3], real_s=[0.8], fake_b=[0.6], fake_s=[0.2],
                fake_ox=np.array([3, 10, 20, 7], dtype=np.int32), fake_oy=np.array([5, 1, 12, 30], dtype=np.int32))
    m.set_input(batch)
    m.optimize_parameters(1, rand=rand)
    torch.cuda.synchronize()
    cfg = O.StepConfig(netG="resnet_9blocks", batch_size_G2=NT, add_fake_T_sample_size=NF, gan_mode=gan_mode)
    sdG, sdD, sdD2 = [copy.deepcopy(s) for s in sds]
    res = O.train_step(cfg, sdG, sdD, sdD2, {}, O.step_inputs_from_batch(batch), rand, step=1)
    losses = m.current_losses()
    for k, v in res["losses"].items():
        assert abs(losses[k] - v) <= GATE * max(1.0, abs(v)), (k, losses[k], v)
    for nm in ("fake_I", "fake_T", "fake_N", "aug_fake_I"):
        assert rel(getattr(m, nm), res[nm]) < GATE, nm
    assert rel(m.pred_fake_T_full.permute(0, 3, 1, 2), res["pred_fake_T_full"]) < GATE
    for tag, net, grads in (("G", m.netG, res["grads_G"]), ("D", m.netD, res["grads_D"]), ("D2", m.netD2, res["grads_D2"])):
        print(tag, "worst grad rel err vs oracle", check_grads(net, {k: v.numpy() for k, v in grads.items()}, tag))


def test_cuda_graph_replay_matches_eager_step():
    """The captured-and-replayed train step (CUDA graph, the bench path) must compute what eager launches compute.
    GAN training with beta1 = 0 Adam is chaotic over several steps (double-precision atomics reorder), so the
    comparison is for ONE step from an identical snapshot: eager vs first replay vs second replay."""
    import vts_b200
    from oracle import skit_oracle as O
    S, NT, NF = 64, 8, 4
    batch = O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True)
    rs = np.random.RandomState(3)
    rands = [dict(real_b=[rs.rand()], real_s=[rs.rand()], fake_b=[rs.rand()], fake_s=[rs.rand()],
                  fake_ox=rs.randint(0, S - 32, NF).astype(np.int32), fake_oy=rs.randint(0, S - 32, NF).astype(np.int32))
             for _ in range(3)]
    torch.manual_seed(2)
    m = vts_b200.SinSKITGModel(vts_b200.default_options(batch_size_G2=NT, add_fake_T_sample_size=NF, cuda_graph=True, cuda_graph_warmup=2))
    nets = (m.netG, m.netD, m.netD2)
    for i in range(2):
        m.set_input(batch)
        m.optimize_parameters(1, rand=rands[i])
    assert m._graph is None

    def snapshot():
        return [[t.clone() for t in (n.flat_param, n.exp_avg, n.exp_avg_sq)] + [b.clone() for b in n.buffers()] for n in nets]

    def restore(snap):
        for n, ts in zip(nets, snap):
            for dst, src in zip([n.flat_param, n.exp_avg, n.exp_avg_sq] + list(n.buffers()), ts):
                dst.copy_(src)
            n.refresh_packs()
        m.step_count = 2

    def run():
        m.set_input(batch)
        m.optimize_parameters(1, rand=rands[2])
        torch.cuda.synchronize()
        return m.current_losses(), m.fake_I.clone(), m.fake_T.clone(), [n.flat_param.clone() for n in nets]

    snap = snapshot()
    res_g = run()                      # captures, then replays once
    g = m._graph
    assert g is not None
    restore(snap)
    m._graph, m.opt.cuda_graph = None, False
    res_e = run()                      # the same step, eager launches
    restore(snap)
    m._graph, m.opt.cuda_graph = g, True
    res_r = run()                      # pure replay
    assert m._graph is g
    for other in (res_e, res_r):
        for k in res_g[0]:
            assert abs(res_g[0][k] - other[0][k]) <= 1e-4 * max(1.0, abs(res_g[0][k])), (k, res_g[0][k], other[0][k])
        assert rel(other[1], res_g[1]) < 1e-5 and rel(other[2], res_g[2]) < 1e-5
        for pa, pb in zip(res_g[3], other[3]):
            # beta1 = 0: each weight moves by ~lr whatever the gradient size; only sign-ambiguous (near-zero
            # gradient) entries may differ between two runs of the same arithmetic with reordered atomics
            assert ((pa - pb).abs() > 2e-4).float().mean().item() < 0.02


def test_train_step_with_patchnce_vs_oracle():
    """BASELINE.json configs[2] wiring at a small size: GAN + L1 + patch-L1 + PatchNCE (5 feature layers, netF='sample').
    The PatchNCE term's value and its gradient contribution to G (through the query encoder pass and through fake_I)."""
    import vts_b200
    from oracle import skit_oracle as O
    S, NT, NF, P = 64, 8, 4, 64
    torch.manual_seed(1)
    opt = vts_b200.default_options(batch_size_G2=NT, add_fake_T_sample_size=NF, lambda_NCE=1.0, num_patches=P)
    m = vts_b200.SinSKITGModel(opt)
    sds = [{k: v.detach().cpu().clone() for k, v in net.state_dict().items()} for net in (m.netG, m.netD, m.netD2)]
    batch = O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True)
    rs = np.random.RandomState(5)
    sizes = [m.netG.feature_hw(l, S, S) for l in m.nce_layers]
    rand = dict(real_b=[0.3], real_s=[0.8], fake_b=[0.6], fake_s=[0.2],
                fake_ox=np.array([3, 10, 20, 7], dtype=np.int32), fake_oy=np.array([5, 1, 12, 30], dtype=np.int32),
                nce_ids=[rs.permutation(h * w)[:min(P, h * w)] for h, w in sizes])
    m.set_input(batch)
    m.optimize_parameters(1, rand=rand)
    torch.cuda.synchronize()
    cfg = O.StepConfig(netG="resnet_9blocks", batch_size_G2=NT, add_fake_T_sample_size=NF, lambda_NCE=1.0, num_patches=P)
    sdG, sdD, sdD2 = [copy.deepcopy(s) for s in sds]
    res = O.train_step(cfg, sdG, sdD, sdD2, {}, O.step_inputs_from_batch(batch), rand, step=1)
    losses = m.current_losses()
    assert "NCE" in losses and "NCE" in res["losses"]
    for k, v in res["losses"].items():
        assert abs(losses[k] - v) <= GATE * max(1.0, abs(v)), (k, losses[k], v)
    print("G worst grad rel err vs oracle (with PatchNCE)", check_grads(m.netG, {k: v.numpy() for k, v in res["grads_G"].items()}, "G"))
    # the NCE term must actually move the gradient: compare with a run without it
    cfg0 = O.StepConfig(netG="resnet_9blocks", batch_size_G2=NT, add_fake_T_sample_size=NF)
    sdG0, sdD0, sdD20 = [copy.deepcopy(s) for s in sds]
    res0 = O.train_step(cfg0, sdG0, sdD0, sdD20, {}, O.step_inputs_from_batch(batch), rand, step=1)
    k = "model.12.conv_block.1.weight"
    assert rel(res["grads_G"][k], res0["grads_G"][k]) > 1e-3


def test_train_step_with_lpips_vs_oracle():
    """The reference's default perceptual terms (lambda_G1_lpips 1, lambda_G2_lpips 10): LPIPS-VGG16 on
    (fake_I, real_I) and on the single-channel touch patches, value and gradient contribution to G, against the oracle's
    train step with the same random LPIPS weights.  Then the same step captured and replayed as a CUDA graph."""
    import vts_b200
    from oracle import skit_oracle as O
    S, NT, NF = 64, 8, 4
    torch.manual_seed(2)
    sdL = O.lpips_random_state(11)
    opt = vts_b200.default_options(batch_size_G2=NT, add_fake_T_sample_size=NF, lambda_G1_lpips=1.0, lambda_G2_lpips=10.0,
                                   lpips_state=sdL)
    m = vts_b200.SinSKITGModel(opt)
    sds = [{k: v.detach().cpu().clone() for k, v in net.state_dict().items()} for net in (m.netG, m.netD, m.netD2)]
    batch = O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True)
    rand = dict(real_b=[0.3], real_s=[0.8], fake_b=[0.6], fake_s=[0.2],
                fake_ox=np.array([3, 10, 20, 7], dtype=np.int32), fake_oy=np.array([5, 1, 12, 30], dtype=np.int32))
    m.set_input(batch)
    m.optimize_parameters(1, rand=rand)
    torch.cuda.synchronize()
    cfg = O.StepConfig(netG="resnet_9blocks", batch_size_G2=NT, add_fake_T_sample_size=NF, lambda_G1_lpips=1.0, lambda_G2_lpips=10.0)
    sdG, sdD, sdD2 = [copy.deepcopy(s) for s in sds]
    res = O.train_step(cfg, sdG, sdD, sdD2, {}, O.step_inputs_from_batch(batch), rand, step=1, sdL=sdL)
    losses = m.current_losses()
    assert "G_lpips" in losses and "G2_lpips" in losses and losses["G_lpips"] > 0 and losses["G2_lpips"] > 0
    for k, v in res["losses"].items():
        assert abs(losses[k] - v) <= GATE * max(1.0, abs(v)), (k, losses[k], v)
    print("G worst grad rel err vs oracle (with LPIPS)", check_grads(m.netG, {k: v.numpy() for k, v in res["grads_G"].items()}, "G"))
    cfg0 = O.StepConfig(netG="resnet_9blocks", batch_size_G2=NT, add_fake_T_sample_size=NF)
    sdG0, sdD0, sdD20 = [copy.deepcopy(s) for s in sds]
    res0 = O.train_step(cfg0, sdG0, sdD0, sdD20, {}, O.step_inputs_from_batch(batch), rand, step=1)
    k = "model.%d.weight" % (12 + 9 + 9)     # the 7x7 head: sees the perceptual gradient directly
    assert rel(res["grads_G"][k], res0["grads_G"][k]) > 1e-3
    # graph capture + replay of the LPIPS step (device-side losses stay finite and close to the eager value)
    for _ in range(4):
        m.set_input(batch)
        m.optimize_parameters(1)
    l2 = m.current_losses()
    assert np.isfinite(l2["G_lpips"]) and np.isfinite(l2["G2_lpips"]) and l2["G_lpips"] > 0


def test_train_step_with_patchnce_mlp_vs_oracle():
    """netF = 'mlp_sample' (CUT's default projection head): the per-layer MLP's forward, its weight gradients, the
    gradient it passes back to the generator, and its own Adam update."""
    import vts_b200
    from oracle import skit_oracle as O
    S, NT, NF, P = 64, 8, 4, 64
    torch.manual_seed(1)
    opt = vts_b200.default_options(batch_size_G2=NT, add_fake_T_sample_size=NF, lambda_NCE=1.0, num_patches=P, netF="mlp_sample",
                                   netF_nc=64, nce_layers="0,4,12")
    m = vts_b200.SinSKITGModel(opt)
    chans = [9, 128, 256]
    m.netF.create_mlp(channels=chans, device=m.device)
    m.netF.flatten_parameters()
    with torch.no_grad():
        m.netF.flat_param.mul_(20.0)       # gain-0.02 init gives a nearly flat projection; scale up for a conditioned test
    sds = [{k: v.detach().cpu().clone() for k, v in net.state_dict().items()} for net in (m.netG, m.netD, m.netD2, m.netF)]
    batch = O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True)
    rs = np.random.RandomState(5)
    sizes = [m.netG.feature_hw(l, S, S) for l in m.nce_layers]
    rand = dict(real_b=[0.3], real_s=[0.8], fake_b=[0.6], fake_s=[0.2],
                fake_ox=np.array([3, 10, 20, 7], dtype=np.int32), fake_oy=np.array([5, 1, 12, 30], dtype=np.int32),
                nce_ids=[rs.permutation(h * w)[:min(P, h * w)] for h, w in sizes])
    m.set_input(batch)
    m.optimize_parameters(1, rand=rand)
    torch.cuda.synchronize()
    cfg = O.StepConfig(netG="resnet_9blocks", batch_size_G2=NT, add_fake_T_sample_size=NF, lambda_NCE=1.0, num_patches=P, nce_layers=(0, 4, 12))
    sdG, sdD, sdD2, sdF = [copy.deepcopy(s) for s in sds]
    res = O.train_step(cfg, sdG, sdD, sdD2, {}, O.step_inputs_from_batch(batch), rand, step=1, sdF=sdF)
    losses = m.current_losses()
    for k, v in res["losses"].items():
        assert abs(losses[k] - v) <= GATE * max(1.0, abs(v)), (k, losses[k], v)
    print("G worst grad rel err", check_grads(m.netG, {k: v.numpy() for k, v in res["grads_G"].items()}, "G"))
    print("F worst grad rel err", check_grads(m.netF, {k: v.numpy() for k, v in res["grads_F"].items()}, "F"))
    for k, v in m.netF.state_dict().items():   # Adam moved the MLP weights like the oracle's
        g = res["grads_F"][k].reshape(-1)
        ok = g.abs() > max(1e-7, 0.05 * float(g.pow(2).mean().sqrt()))
        bad = ((v.detach().cpu().reshape(-1) - sdF[k].detach().reshape(-1)).abs() > 2e-4)[ok]
        assert bad.float().mean().item() < 0.02, k


def test_inference_forward_graph_and_batches():
    """test() (generator forward, BASELINE.json configs[0]/[4]) with batch > 1 and CUDA-graph replay vs the oracle."""
    import vts_b200
    from oracle import skit_oracle as O
    torch.manual_seed(4)
    opt = vts_b200.default_options(isTrain=False, ngf=64)
    m = vts_b200.SinSKITGModel(opt)
    sd = {k: v.detach().cpu().clone() for k, v in m.netG.state_dict().items()}
    B, S = 3, 64
    g = torch.Generator().manual_seed(9)
    batch = {"S": torch.rand(B, 1, S, S, generator=g) * 2 - 1, "M": (torch.rand(B, 1, S, S, generator=g) > 0.1).float()}
    pe = O.spe_grid(S, S, 4, B)
    ref = O.model_forward(O.StepConfig(), sd, batch["S"] * batch["M"], pe, batch["M"])
    for i in range(3):          # eager, capture, replay
        m.set_input(batch, phase="test")
        fI, fT, fN = m.test()
        torch.cuda.synchronize()
        assert rel(fI, ref["fake_I"]) < GATE and rel(fT, ref["fake_T"]) < GATE and rel(fN, ref["fake_N"]) < GATE, i
    assert m._tgraph is not None


def test_unsupported_gan_modes_raise():
    """'lsgan' / 'vanilla' / 'wgan(gp)' make GANLoss return a 0-d tensor, on which the reference's own compute_G2_loss fails
    (len() of a 0-d tensor, sinskitG_model.py:1782-1784): the B200 step refuses them instead of silently running another loss."""
    import vts_b200
    for mode in ("lsgan", "vanilla", "wgangp"):
        with pytest.raises(NotImplementedError, match="gan_mode"):
            vts_b200.SinSKITGModel(vts_b200.default_options(gan_mode=mode, batch_size_G2=8, add_fake_T_sample_size=4))
