"""Pin the oracle (oracle/skit_oracle.py) against fixtures generated from the REAL reference
code by oracle/make_golden.py (the reference ships no golden vectors)."""
import math
import os
import random

import numpy as np
import pytest
import torch

from oracle import skit_oracle as O

TOL = dict(rtol=1e-4, atol=2e-5)


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name), allow_pickle=False))


def sd_from(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(v.copy()) for k, v in z.items() if k.startswith(prefix)}


def rand_input(seed, *shape):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * 2 - 1


def close(a, b, **kw):
    tol = dict(TOL)
    tol.update(kw)
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), **tol)


def sub(t, s):
    return t[..., ::s, ::s].contiguous().numpy()


def test_resnet_generator(golden_dir):
    z = load(golden_dir, "networks.npz")
    sd = sd_from(z, "Gres.")
    x = rand_input(11, 1, 9, 48, 40)
    y = O.resnet_g_forward(sd, x)
    close(y, z["Gres_out"])
    y2, feats = O.resnet_g_forward(sd, x, layers=[0, 4, 8, 12, 16])
    close(y2, z["Gres_out"])
    for i, f in enumerate(feats):
        close(sub(f, 2), z["Gres_feat%d" % i])
        assert abs(f.double().norm().item() / z["Gres_feat%d_norm" % i] - 1) < 1e-5
    enc = O.resnet_g_forward(sd, x, layers=[0, 4, 8], encode_only=True)
    assert len(enc) == 3


def test_unet_generator(golden_dir):
    z = load(golden_dir, "networks.npz")
    x = rand_input(12, 1, 9, 256, 256)
    y = O.unet_custom_forward(sd_from(z, "Gunet."), x)
    close(sub(y, 4), z["Gunet_out"])
    assert abs(y.double().norm().item() / z["Gunet_out_norm"] - 1) < 1e-5
    ys = O.unet_custom_forward(sd_from(z, "Gunet_style."), x, style_code=torch.from_numpy(z["Gunet_style_code"]))
    close(sub(ys, 4), z["Gunet_style_out"])


@pytest.mark.parametrize("tag,n_blocks,res,batch", [("sg2", 6, 128, 1), ("sg2small", 2, 64, 2)])
def test_stylegan2_generator(golden_dir, tag, n_blocks, res, batch):
    z = load(golden_dir, "stylegan2.npz")
    sd = sd_from(z, tag + ".")
    x = rand_input(22, batch, 9, res, res)
    noises = [torch.from_numpy(z[tag + "_noise"])] if tag == "sg2" else None   # 'small' variant: no noise injection
    y, feats = O.stylegan2_g_forward(sd, x, n_blocks=n_blocks, layers=[1, 2, 3], noises=noises)
    close(y, z[tag + "_out"])
    assert len(feats) == 3
    for i, f in enumerate(feats):
        close(sub(f, 4), z[tag + "_feat%d" % i])
        assert abs(f.double().norm().item() / z[tag + "_feat%d_norm" % i] - 1) < 1e-5
    assert len(O.stylegan2_g_forward(sd, x, n_blocks=n_blocks, layers=[1, 2], encode_only=True)) == 2


def test_lpips_trunk_matches_torchvision_vgg16():
    """The oracle's VGG16 feature slices against torchvision's vgg16 module carrying the same weights (the LPIPS package
    wraps exactly that module); then the LPIPS value is checked for its defining properties."""
    tv = pytest.importorskip("torchvision")
    sd = O.lpips_random_state(3)
    net = tv.models.vgg16(weights=None).features.eval()
    with torch.no_grad():
        for s, idxs in enumerate(O.VGG_SLICES):
            for i in idxs:
                net[i].weight.copy_(sd["net.slice%d.%d.weight" % (s + 1, i)])
                net[i].bias.copy_(sd["net.slice%d.%d.bias" % (s + 1, i)])
    x = rand_input(31, 2, 3, 32, 48)
    feats = O.vgg16_features(sd, x)
    ends = (4, 9, 16, 23, 30)
    h = x
    with torch.no_grad():
        for k, e in enumerate(ends):
            h = net[(0 if k == 0 else ends[k - 1]):e](h)
            close(feats[k], h)
    y = rand_input(32, 2, 3, 32, 48)
    v = O.lpips_vgg(sd, x, y)
    assert v.shape == (2, 1, 1, 1) and (v > 0).all()
    assert O.lpips_vgg(sd, x, x).abs().max() == 0
    close(O.lpips_vgg(sd, x, y), O.lpips_vgg(sd, y, x))                       # symmetric
    close(O.lpips_vgg(sd, x[:, :1], y[:, :1]), O.lpips_vgg(sd, x[:, :1].repeat(1, 3, 1, 1), y[:, :1].repeat(1, 3, 1, 1)))


def test_multiscale_discriminator_and_ganloss(golden_dir):
    z = load(golden_dir, "networks.npz")
    sd = sd_from(z, "D_before.")
    pred = O.multiscale_d_forward(sd, rand_input(14, 6, 7, 32, 32))
    for i, p in enumerate(pred):
        close(p[-1], z["D_pred%d" % i])
    close(O.gan_loss(pred, False), z["D_loss_fake"])
    close(O.gan_loss(pred, True), z["D_loss_real"])
    close(O.gan_loss(pred[0][-1], True), z["D_loss_tensor_real"])
    for k, v in sd_from(z, "D_after.").items():
        close(sd[k], v)
    sdb = sd_from(z, "Dbasic.")
    close(O.nlayer_d_forward(sdb, "model.", rand_input(15, 1, 4, 70, 58)), z["Dbasic_pred"])


def test_blur_resamplers(golden_dir):
    z = load(golden_dir, "networks.npz")
    xr = rand_input(16, 2, 3, 10, 14)
    close(O.blur_down(xr), z["blur_down"])
    close(O.blur_up(xr), z["blur_up"])


def test_spe(golden_dir):
    z = load(golden_dir, "ops.npz")
    close(O.spe_grid(20, 28, 4, 2), z["spe_20x28"], rtol=1e-5, atol=1e-6)
    close(O.spe_grid(1100, 8, 4, 1), z["spe_1100"], rtol=1e-5, atol=1e-5)


def test_patch_gather(golden_dir):
    z = load(golden_dir, "ops.npz")
    img = rand_input(21, 1, 3, 96, 80)
    ox, oy, cs = O.patch_offsets_from_coords(z["gather_coords"])
    assert np.array_equal(ox, z["gather_ox"].reshape(-1)) and np.array_equal(oy, z["gather_oy"].reshape(-1))
    assert np.array_equal(cs, z["gather_cs"].reshape(-1))
    out = O.get_patch_in_input(img, z["gather_coords"])
    assert np.array_equal(out.numpy(), z["gather_out"])  # pure copy: bit exact
    M = torch.from_numpy(z["rand_M"])
    img2 = rand_input(22, 1, 2, 96, 96)
    random.seed(5)
    smp, rox, roy, _ = O.get_patch_in_input(img2, None, sample_size=12, M=M, return_offset=True)
    assert np.array_equal(rox, z["rand_ox"]) and np.array_equal(roy, z["rand_oy"])
    assert np.array_equal(smp.numpy(), z["rand_out"])
    again = O.get_patch_in_input(img, None, sample_size=12, offset_x=rox, offset_y=roy)
    assert np.array_equal(again.numpy(), z["rand_from_offsets"])


def test_normal_and_diffaugment(golden_dir):
    z = load(golden_dir, "ops.npz")
    T = rand_input(23, 3, 2, 9, 7)
    T[0, :, 0, 0] = 0
    close(O.compute_normal(T, 0.25), z["normal_025"], rtol=1e-6, atol=1e-7)
    close(O.compute_normal(T, 0.0), z["normal_0"], rtol=1e-6, atol=1e-7)
    xa = rand_input(24, 2, 3, 12, 10)
    u = z["diffaug_u"]
    close(O.diffaugment_bs(xa, u[0], u[1]), z["diffaug_bs"], rtol=1e-6, atol=1e-6)


def test_patchsample_and_patchnce(golden_dir):
    z = load(golden_dir, "ops.npz")
    feats = [rand_input(25, 2, 6, 8, 8), rand_input(26, 2, 12, 4, 4)]
    ids = [z["psf_ids0"], z["psf_ids1"]]
    o = O.patch_sample_f(feats, ids)
    close(o[0], z["psf_out0"], rtol=1e-6, atol=1e-6)
    close(o[1], z["psf_out1"], rtol=1e-6, atol=1e-6)
    mlps = [tuple(torch.from_numpy(z["psf_mlp.mlp_%d.%d.%s" % (i, j, w)]) for j in (0, 2) for w in ("weight", "bias"))
            for i in range(2)]
    om = O.patch_sample_f(feats, ids, mlps)
    close(om[0], z["psf_mlp_out0"])
    close(om[1], z["psf_mlp_out1"])
    for nm, allneg in (("nce_same", False), ("nce_all", True)):
        q = torch.from_numpy(z[nm + "_q"]).requires_grad_(True)
        k = torch.from_numpy(z[nm + "_k"])
        loss = O.patchnce_loss(q, k, 0.07, batch_size=2, all_negatives_from_minibatch=allneg)
        close(loss.detach(), z[nm + "_loss"])
        loss.mean().backward()
        close(q.grad, z[nm + "_dq"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("tag,netG", [("resnet", "resnet_9blocks"), ("unet", "unet256_custom")])
def test_train_step_matches_reference(golden_dir, tag, netG):
    """Full optimize_parameters: losses, outputs, every gradient, post-Adam weights, BN stats."""
    z = load(golden_dir, "step_%s.npz" % tag)
    S, NT, NF = [int(v) for v in z["meta"]]
    cfg = O.StepConfig(netG=netG, batch_size_G2=NT, add_fake_T_sample_size=NF)
    sdG, sdD, sdD2 = sd_from(z, "G_before."), sd_from(z, "D_before."), sd_from(z, "D2_before.")
    batch = O.step_inputs_from_batch(O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True))
    u = z["rand_u"]
    rand = dict(real_b=u[0], real_s=u[1], fake_b=u[2], fake_s=u[3], fake_ox=z["fake_ox"], fake_oy=z["fake_oy"])
    res = O.train_step(cfg, sdG, sdD, sdD2, {}, batch, rand, step=1)
    for k, v in res["losses"].items():
        assert abs(v - float(z["loss_l_" + k])) <= 2e-5 * max(1.0, abs(v)), (k, v, float(z["loss_l_" + k]))
    for nm in ("fake_I", "fake_T", "fake_N", "aug_fake_I"):
        close(sub(res[nm], 2), z[nm], rtol=1e-4, atol=1e-6)
    close(res["pred_fake_T_full"], z["pred_fake_T_full"], rtol=1e-4, atol=1e-6)
    for net, grads, sd in (("G", res["grads_G"], sdG), ("D", res["grads_D"], sdD), ("D2", res["grads_D2"], sdD2)):
        keys = [k[len(net) + 6:] for k in z if k.startswith(net + "_grad.")]
        assert keys
        for k in keys:
            g_ref = z["%s_grad.%s" % (net, k)]
            assert k in grads, k
            err = np.linalg.norm(grads[k].numpy() - g_ref) / max(np.linalg.norm(g_ref), 1e-30)
            # a conv bias feeding a norm layer has a mathematically zero gradient: what the
            # reference holds there is rounding noise, so judge it on the weight-grad scale
            wk = "%s_grad.%s" % (net, k.replace(".bias", ".weight"))
            noise = k.endswith(".bias") and wk in z and np.linalg.norm(g_ref) < 1e-3 * np.linalg.norm(z[wk])
            assert err < 2e-3 or noise, (net, k, err)
        for k, v in sd.items():
            sk = "%s_after_sub.%s" % (net, k)
            gk = "%s_grad.%s" % (net, k)
            if sk in z and gk in z:
                # post-Adam weights (strided subsample).  First step with beta1=0 moves every weight
                # by ~lr*sign(g); skip the pure-noise-gradient biases and near-zero-gradient elements.
                g = z[gk].reshape(-1)[::7]
                ok = np.abs(g) > max(1e-6, 0.05 * float(np.sqrt(np.mean(z[gk] ** 2))))
                wk = "%s_grad.%s" % (net, k.replace(".bias", ".weight"))
                if k.endswith(".bias") and wk in z and np.linalg.norm(z[gk]) < 1e-3 * np.linalg.norm(z[wk]):
                    continue
                close(v.reshape(-1)[::7].numpy()[ok], z[sk][ok], rtol=1e-4, atol=2e-6)
            if ("%s_after.%s" % (net, k)) in z:
                # running means inherit +-lr*0.1 of noise from the zero-gradient conv biases that Adam
                    # moved by +-lr in the D step (sign of rounding noise) -> absolute tolerance 3e-4
                    close(v, z["%s_after.%s" % (net, k)], rtol=1e-4, atol=3e-4)


def test_train_step_lpips_wiring_matches_reference(golden_dir):
    """The oracle's LPIPS wiring (full-image term, per-channel touch-patch term, reductions, weights, backward into G)
    against the REAL reference's optimize_parameters run with the same criterion plugged into its `lpips.LPIPS` call sites
    (tests/golden/step_resnet_lpips.npz, oracle/make_golden.py: make_step_lpips).  Initial weights: step_resnet.npz."""
    z0 = load(golden_dir, "step_resnet.npz")
    z = load(golden_dir, "step_resnet_lpips.npz")
    S, NT, NF, lp_seed = [int(v) for v in z["meta"]]
    sdG, sdD, sdD2 = sd_from(z0, "G_before."), sd_from(z0, "D_before."), sd_from(z0, "D2_before.")
    norm = sum(v.double().pow(2).sum().item() for v in sdG.values()) ** 0.5
    assert abs(norm / float(z["G_before_norm"]) - 1) < 1e-6          # same initial generator as the fixture's run
    cfg = O.StepConfig(netG="resnet_9blocks", batch_size_G2=NT, add_fake_T_sample_size=NF, lambda_G1_lpips=1.0, lambda_G2_lpips=10.0)
    batch = O.step_inputs_from_batch(O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True))
    u = z["rand_u"]
    rand = dict(real_b=u[0], real_s=u[1], fake_b=u[2], fake_s=u[3], fake_ox=z["fake_ox"], fake_oy=z["fake_oy"])
    res = O.train_step(cfg, sdG, sdD, sdD2, {}, batch, rand, step=1, sdL=O.lpips_random_state(lp_seed))
    assert "G_lpips" in res["losses"] and "G2_lpips" in res["losses"]
    for k, v in res["losses"].items():
        ref = float(z["loss_l_" + k])
        assert abs(v - ref) <= 2e-5 * max(1.0, abs(v)), (k, v, ref)
    keys = [k[len("G_grad."):] for k in z if k.startswith("G_grad.")]
    assert len(keys) > 40
    for k in keys:
        g_ref = z["G_grad." + k]
        err = np.linalg.norm(res["grads_G"][k].numpy() - g_ref) / max(np.linalg.norm(g_ref), 1e-30)
        wk = "G_grad." + k.replace(".bias", ".weight")
        noise = k.endswith(".bias") and wk in z and np.linalg.norm(g_ref) < 1e-3 * np.linalg.norm(z[wk])
        assert err < 2e-3 or noise, (k, err)
    # the perceptual terms moved the generator's gradient (they are not a no-op in this fixture)
    k = "model.30.weight"
    assert np.linalg.norm(z["G_grad." + k] - z0["G_grad." + k]) > 1e-3 * np.linalg.norm(z0["G_grad." + k])


def test_standalone_init_tables_match_reference_keys_and_shapes(golden_dir):
    """O.init_resnet_g / O.init_multiscale_d (the weights bench.py's reference arm starts from, built without the product
    package) carry exactly the reference networks' state_dict keys, shapes and blur buffers, and its xavier(0.02) statistics."""
    g = np.load(os.path.join(golden_dir, "networks.npz"))
    gu = np.load(os.path.join(golden_dir, "step_unet.npz"))
    ref_u = {k[len("G_before."):]: gu[k] for k in gu.files if k.startswith("G_before.")}
    sd_u = O.init_unet_custom(9, 4, 8, 4, seed=3)        # the reference's default generator (unet256_custom) at the fixture's ngf 4
    assert set(sd_u) == set(ref_u) and all(tuple(v.shape) == ref_u[k].shape for k, v in sd_u.items())
    for prefix, sd in (("Gres.", O.init_resnet_g(9, 5, 8, 9, seed=3)), ("D_before.", O.init_multiscale_d(7, 8, 3, 3, seed=3))):
        ref = {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}
        assert set(sd) == set(ref), prefix
        for k, v in sd.items():
            assert tuple(v.shape) == ref[k].shape, (prefix, k)
            if k.endswith(".filt"):
                np.testing.assert_allclose(v.numpy(), ref[k], rtol=1e-6)
    sd = O.init_resnet_g(9, 5, 64, 9, seed=0)
    w = sd["model.12.conv_block.1.weight"]
    assert abs(float(w.std()) - 0.02 * math.sqrt(2.0 / (2 * 256 * 9))) < 2e-6 and float(sd["model.12.conv_block.1.bias"].abs().max()) == 0.0
    y = O.resnet_g_forward(sd, torch.rand(1, 9, 32, 32) * 2 - 1)
    assert y.shape == (1, 5, 32, 32) and torch.isfinite(y).all()


def test_evaluation_metrics_oracle(golden_dir):
    """T_AE / T_MSE restatements against the reference's own functions (tests/golden/metrics.npz); the torchmetrics restatements
    (PSNR, SSIM: parity unpinned, the package is absent) against their defining properties."""
    z = np.load(os.path.join(golden_dir, "metrics.npz"))
    rT, fT = torch.from_numpy(z["real_T"]), torch.from_numpy(z["fake_T"])
    assert abs(O.normal_angle_error_deg(rT, fT.clamp(0, 1), 1.0).item() - float(z["T_AE"])) < 1e-4
    g = torch.Generator().manual_seed(1)
    a = torch.rand(2, 3, 40, 48, generator=g)
    b = (a + 0.05 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    assert abs(O.ssim_torchmetrics(a, a).item() - 1.0) < 1e-6 and 0.3 < O.ssim_torchmetrics(a, b).item() < 0.999
    assert abs(O.ssim_torchmetrics(a, b).item() - O.ssim_torchmetrics(b, a).item()) < 1e-6          # symmetric
    assert abs(O.psnr_torchmetrics(a, b).item() - 10 * math.log10(1.0 / ((a - b) ** 2).mean().item())) < 1e-4
    m = O.evaluation_metrics(a * 2 - 1, b * 2 - 1, rT, fT)
    assert abs(m["T_MSE"] - float(z["T_MSE"])) < 1e-7 and abs(m["T_AE"] - float(z["T_AE"])) < 1e-4
    # the min-max rescale uses the REAL image's range (model_utils.py:483-487): an affine change of both images is undone
    m2 = O.evaluation_metrics((a * 2 - 1) * 3 + 1, (b * 2 - 1) * 3 + 1, rT, fT)
    assert abs(m2["I_PSNR"] - m["I_PSNR"]) < 1e-3 and abs(m2["I_SSIM"] - m["I_SSIM"]) < 1e-5


def test_gan_loss_oracle_all_modes(golden_dir):
    """O.gan_loss in every mode of the reference's GANLoss (values and autograd gradients) against tests/golden/ganloss.npz."""
    g = np.load(os.path.join(golden_dir, "ganloss.npz"))
    preds = [torch.from_numpy(g["pred%d" % i]) for i in range(3)]
    for mode in ("nonsaturating", "hinge", "wgan", "wgangp", "lsgan", "vanilla"):
        for smooth in (0, 1):
            for is_real in (1, 0):
                key = "%s/%d/%d" % (mode, smooth, is_real)
                ps = [p.clone().requires_grad_(True) for p in preds]
                loss = O.gan_loss([[p] for p in ps], bool(is_real), mode, real_label=0.8 if smooth else 1.0, fake_label=0.0)
                loss.mean().backward()
                np.testing.assert_allclose(loss.detach().numpy(), g[key + "/multi"], rtol=1e-6, atol=1e-7, err_msg=key)
                for i, p in enumerate(ps):
                    np.testing.assert_allclose(p.grad.numpy(), g[key + "/grad%d" % i], rtol=1e-6, atol=1e-8, err_msg=key)
                bare = O.gan_loss([preds[0]], bool(is_real), mode, real_label=0.8 if smooth else 1.0, fake_label=0.0)
                np.testing.assert_allclose(bare.numpy(), g[key + "/bare"], rtol=1e-6, atol=1e-7, err_msg=key)
