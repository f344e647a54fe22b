"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the REAL reference code.

Run in the build container only (needs /root/reference):

    cd /root/repo && python oracle/make_golden.py

Every fixture stores the weights, the seed recipe for the inputs (torch CPU generator,
platform-stable) and the outputs of the unmodified reference functions.  Large tensors are
stored as a strided subsample plus their L2 norm so the committed files stay small.
"""
import argparse
import contextlib
import io
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference, ref_networks  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def sub(t, stride=4):
    """strided subsample of the two trailing dims + L2 norm, as float32 numpy."""
    t = t.detach().float()
    return t[..., ::stride, ::stride].contiguous().numpy(), np.float64(t.double().norm().item())


def sd_np(sd, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def rand_input(seed, *shape):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * 2 - 1


def make_networks():
    N = ref_networks()
    opt = argparse.Namespace(batch_size=1, T_resolution_multiplier=1, gan_mode="nonsaturating")
    out = {}
    # ---- resnet generator (a3)
    torch.manual_seed(1)
    with quiet():
        G = N.define_G(9, 5, 8, "resnet_9blocks", "instance", False, "xavier", 0.02, False, False, [], opt,
                       num_layer_separate=4)
    # biases are zero-initialised; perturb them so bias handling is exercised
    with torch.no_grad():
        for k, p in G.named_parameters():
            if k.endswith("bias"):
                p.normal_(0, 0.02)
            else:
                p.mul_(20.0)  # gain 0.02 init gives tiny outputs; scale up for a better conditioned test
    x = rand_input(11, 1, 9, 48, 40)
    with torch.no_grad(), quiet():
        y = G(x)
        yf, feats = G(x, layers=[0, 4, 8, 12, 16])
    out.update(sd_np(G.state_dict(), "Gres."))
    out["Gres_out"] = y.numpy()
    for i, f in enumerate(feats):
        out["Gres_feat%d" % i], out["Gres_feat%d_norm" % i] = sub(f, 2)
    # ---- default custom U-Net generator (a2)
    torch.manual_seed(2)
    with quiet():
        U = N.define_G(9, 5, 4, "unet256_custom", "instance", False, "xavier", 0.02, False, False, [], opt,
                       num_layer_separate=4)
    with torch.no_grad():
        for k, p in U.named_parameters():
            if k.endswith("bias"):
                p.normal_(0, 0.05)
            else:
                p.mul_(20.0)
    x = rand_input(12, 1, 9, 256, 256)
    with torch.no_grad():
        y = U(x)
    out.update(sd_np(U.state_dict(), "Gunet."))
    out["Gunet_out"], out["Gunet_out_norm"] = sub(y, 4)
    # ---- skitG variant: style code concat/tile at the innermost level (a2, skitG_model.py:1294-1302)
    opt_s = argparse.Namespace(batch_size=1, T_resolution_multiplier=1, gan_mode="nonsaturating",
                               use_style_code=True, style_code_mode="concat", style_code_mapping_mode="tile",
                               style_code_dim=16, num_layer_style_code=1)
    torch.manual_seed(3)
    with quiet():
        Us = N.CustomUnetGenerator(9, 5, num_downs=8, ngf=4, norm_layer=N.get_norm_layer("instance"),
                                   num_layer_separate=4, opt=opt_s, input_size=256)
        N.init_weights(Us, "xavier", 0.02)
    with torch.no_grad():
        for k, p in Us.named_parameters():
            if "style_code_mapping" in k:
                continue
            p.mul_(20.0) if not k.endswith("bias") else p.normal_(0, 0.05)
    sc = torch.randn(1, 16, generator=torch.Generator().manual_seed(13)).half()
    with torch.no_grad():
        y = Us(x, style_code=sc)
    out.update({("Gunet_style." + k): v.numpy() for k, v in Us.state_dict().items() if "style_code_mapping" not in k})
    out["Gunet_style_code"] = sc.float().numpy()
    out["Gunet_style_out"], out["Gunet_style_out_norm"] = sub(y, 4)
    # ---- multiscale PatchGAN (a5) + GANLoss (a6)
    torch.manual_seed(4)
    with quiet():
        D = N.define_D(7, 8, "multiscale", 3, "batch", "xavier", 0.02, False, num_D=3, gpu_ids=[], opt=opt)
    with torch.no_grad():
        for k, p in D.named_parameters():
            if p.dim() == 4:
                p.mul_(20.0)
            elif k.endswith("bias"):
                p.normal_(0, 0.05)
    out.update(sd_np(D.state_dict(), "D_before."))
    xd = rand_input(14, 6, 7, 32, 32)
    D.train()
    pred = D(xd)
    crit = N.GANLoss("nonsaturating")
    for i, p in enumerate(pred):
        out["D_pred%d" % i] = p[-1].detach().numpy()
    out["D_loss_fake"] = crit(pred, False).detach().numpy()
    out["D_loss_real"] = crit(pred, True).detach().numpy()
    out["D_loss_tensor_real"] = crit(pred[0][-1], True).detach().numpy()  # bare tensor: last batch element quirk
    out.update({("D_after." + k): v.numpy() for k, v in D.state_dict().items() if "running" in k or "num_batches" in k})
    # 'basic' (single NLayer) discriminator on a non-square image
    torch.manual_seed(5)
    with quiet():
        Db = N.define_D(4, 8, "basic", 3, "batch", "xavier", 0.02, False, gpu_ids=[], opt=opt)
    with torch.no_grad():
        for k, p in Db.named_parameters():
            if p.dim() == 4:
                p.mul_(20.0)
    out.update(sd_np(Db.state_dict(), "Dbasic."))
    xb = rand_input(15, 1, 4, 70, 58)
    Db.train()
    out["Dbasic_pred"] = Db(xb).detach().numpy()
    # ---- blur resamplers (A.3)
    xr = rand_input(16, 2, 3, 10, 14)
    out["blur_down"] = N.Downsample(3)(xr).numpy()
    out["blur_up"] = N.Upsample(3)(xr).numpy()
    np.savez_compressed(os.path.join(OUT, "networks.npz"), **out)
    print("networks.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


def make_step_lpips(S=64, NT=8, NF=4):
    """The REAL reference's optimize_parameters with its LPIPS terms on (lambda_G1_lpips 1, lambda_G2_lpips 10, the parser's
    defaults).  The pip package `lpips` is absent, so `lpips.LPIPS` is replaced by a module that evaluates the oracle's
    restatement (skit_oracle.lpips_vgg) with seeded random weights: everything AROUND the criterion — the reference's call
    sites, the per-channel touch-patch form, the view/sum/mean reductions, the loss weights and the backward into G — is the
    reference's own code (sinskitG_model.py:1619-1658, 1707-1715, 1826-1840).  Same seeds and architecture as
    make_step('resnet', 64, 8, 4, ...): the initial weights are those of step_resnet.npz, so only the results are stored."""
    load_reference()
    import torch.nn as nn
    sys.path.insert(0, ROOT)
    from oracle import skit_oracle as O

    class OracleLPIPS(nn.Module):
        def __init__(self, net="vgg", **kw):
            super().__init__()
            self.sd = O.lpips_random_state(11)

        def forward(self, a, b):
            return O.lpips_vgg(self.sd, a, b)

    cwd = os.getcwd()
    os.chdir(os.environ.get("VTS_REFERENCE_ROOT", "/root/reference"))
    old = sys.modules["lpips"].LPIPS
    sys.modules["lpips"].LPIPS = OracleLPIPS
    try:
        with quiet():
            from options.train_options import TrainOptions
            from models import create_model
        cmd = ("--model sinskitG --gpu_ids -1 --name golden --checkpoints_dir /tmp/vts_golden "
               "--crop_size %d --center_w %d --center_h %d --lambda_G1_lpips 1 --lambda_G2_lpips 10 "
               "--use_vision_aided_loss False --batch_size_G2 %d --add_fake_T_sample_size %d "
               "--netG resnet_9blocks --ngf 8" % (S, S, S, NT, NF)).split()
        torch.manual_seed(100)
        with quiet():
            to = TrainOptions()
            to.cmd_line = cmd
            opt = to.parse()
            model = create_model(opt)
            model.setup(opt)
        model.train()
        out = {"G_before_norm": np.float64(sum(v.double().pow(2).sum().item() for v in model.netG.state_dict().values()) ** 0.5)}
        batch = O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True)
        torch.manual_seed(200)
        random.seed(300)
        with quiet():
            model.set_input(batch, phase="train")
            model.optimize_parameters(1)
        torch.manual_seed(200)
        out["rand_u"] = np.array([torch.rand(1, 1, 1, 1).item() for _ in range(4)], dtype=np.float32)
        out["fake_ox"] = model.fake_sample_offset_x.numpy().reshape(-1).astype(np.int32)
        out["fake_oy"] = model.fake_sample_offset_y.numpy().reshape(-1).astype(np.int32)
        for k, v in model.get_current_losses().items():
            out["loss_" + k] = np.float64(v)
        for k, p in model.netG.named_parameters():
            if p.grad is not None:
                out["G_grad." + k] = p.grad.numpy()
        out["meta"] = np.array([S, NT, NF, 11])
    finally:
        sys.modules["lpips"].LPIPS = old
        os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "step_resnet_lpips.npz"), **out)
    print("step_resnet_lpips.npz:", len(out), "arrays;", {k: float(v) for k, v in out.items() if k.startswith("loss_") and "lpips" in k})


def make_options():
    """Defaults of the reference's own option parser (TrainOptions + sinskitG model options) as JSON."""
    import json
    load_reference()
    cwd = os.getcwd()
    os.chdir(os.environ.get("VTS_REFERENCE_ROOT", "/root/reference"))
    try:
        with quiet():
            from options.train_options import TrainOptions
            to = TrainOptions()
            to.cmd_line = "--model sinskitG --gpu_ids -1 --name golden --checkpoints_dir /tmp/vts_golden".split()
            opt = to.parse()
    finally:
        os.chdir(cwd)
    d = {k: v for k, v in vars(opt).items() if isinstance(v, (int, float, str, bool, list, type(None)))}
    with open(os.path.join(OUT, "options.json"), "w") as f:
        json.dump(d, f, indent=0, sort_keys=True)
    print("options.json:", len(d), "options")


def make_stylegan2():
    """StyleGAN2Generator (a4) through the reference's define_G.  ModulatedConv2d builds its unit style with `.cuda()`
    (stylegan_networks.py:310); on this CPU-only container Tensor.cuda is patched to the identity for the duration of the
    call — the reference source is untouched."""
    N = ref_networks()
    out = {}
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for tag, netG, res, ngf, batch in (("sg2", "stylegan2", 128, 4, 1), ("sg2small", "smallstylegan2", 64, 4, 2)):
            opt = argparse.Namespace(load_size=res, crop_size=res, stylegan2_G_num_downsampling=1, netG=netG)
            torch.manual_seed(21)
            with quiet():
                G = N.define_G(9, 5, ngf, netG, "instance", False, "xavier", 0.02, False, False, [], opt)
            with torch.no_grad():   # biases and the noise strength start at zero: perturb so they are exercised
                for k, p in G.named_parameters():
                    if k.endswith("bias"):
                        p.normal_(0, 0.3)
                    elif k.endswith("noise.weight"):
                        p.fill_(0.37)
            x = rand_input(22, batch, 9, res, res)
            # NoiseInjection draws image.new_empty(n,1,H,W).normal_() from the global CPU generator: same seed, same draw
            torch.manual_seed(23)
            noise = torch.empty(batch, 1, res, res).normal_()
            torch.manual_seed(23)
            with torch.no_grad():
                y, feats = G(x, layers=[1, 2, 3])
            out.update(sd_np(G.state_dict(), tag + "."))
            out[tag + "_noise"] = noise.numpy()
            out[tag + "_out"] = y.numpy()
            for i, f in enumerate(feats):
                out[tag + "_feat%d" % i], out[tag + "_feat%d_norm" % i] = sub(f, 4)
    finally:
        torch.Tensor.cuda = orig_cuda
    np.savez_compressed(os.path.join(OUT, "stylegan2.npz"), **out)
    print("stylegan2.npz:", len(out), "arrays")


def make_ops():
    load_reference()
    with quiet():
        from models.model_utils import get_patch_in_input, compute_normal, find_coords_for_patch
        from models.patchnce import PatchNCELoss
        from models.networks import PatchSampleF
        from thirdparty.DiffAugment import DiffAugment
        from thirdparty.mmgeneration.positional_encoding import SinusoidalPositionalEmbedding as SPE
    out = {}
    # SPE (a16)
    out["spe_20x28"] = SPE(4, 0, 1024)(torch.zeros(2, 1, 20, 28)).numpy()
    out["spe_1100"], out["spe_1100_norm"] = sub(SPE(4, 0, 1024)(torch.zeros(1, 1, 1100, 8)), 1)
    # patch gather (a8): known coords incl. border-clamped and half-integer rounding cases
    img = rand_input(21, 1, 3, 96, 80)
    coords = np.zeros((1, 10, 8), dtype=np.float64)
    rs = np.random.RandomState(3)
    coords[0, :, 0] = rs.randint(-10, 70, 10)
    coords[0, :, 1] = rs.randint(-10, 90, 10)
    coords[0, :, 2:4] = 64
    coords[0, :, 4] = 32
    coords[0, :, 5] = 1
    coords[0, :, 6] = rs.randint(0, 32, 10) + 0.5  # exercises np.round half-to-even
    coords[0, :, 7] = rs.randint(0, 32, 10)
    out["gather_coords"] = coords
    out["gather_out"] = get_patch_in_input(img, coords).numpy()
    ox, oy, cs = find_coords_for_patch(coords)
    out["gather_ox"], out["gather_oy"], out["gather_cs"] = ox.numpy(), oy.numpy(), cs.numpy()
    # random mode with an elliptical mask (erosion sampler) -- python `random` stream
    S = 96
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    M = ((((yy - S / 2) / (0.45 * S)) ** 2 + ((xx - S / 2) / (0.40 * S)) ** 2) <= 1).float()[None, None]
    img2 = rand_input(22, 1, 2, S, S)
    one = lambda v: torch.tensor([v])
    aug = dict(H=one(S), W=one(S), scale_factor_h=one(1), scale_factor_w=one(1), crop_pos_x=one(0), crop_pos_y=one(0),
               resize_ratio_w=one(1.0), resize_ratio_h=one(1.0))
    random.seed(5)
    smp, rox, roy, rcs = get_patch_in_input(img2, coords=None, sample_size=12, return_offset=True, center_h=64,
                                            center_w=64, augmentation_params=aug, M=M, device="cpu")
    out["rand_M"] = M.numpy()
    out["rand_out"] = smp.numpy()
    out["rand_ox"], out["rand_oy"] = rox.numpy().reshape(-1), roy.numpy().reshape(-1)
    out["rand_from_offsets"] = get_patch_in_input(img, coords=None, sample_size=12, offset_x=rox, offset_y=roy,
                                                  cutout_size=rcs).numpy()
    # normal (K16), DiffAugment (a15)
    T = rand_input(23, 3, 2, 9, 7)
    T[0, :, 0, 0] = 0
    out["normal_025"] = compute_normal(T, scale_nz=0.25).numpy()
    out["normal_0"] = compute_normal(T, scale_nz=0).numpy()
    xa = rand_input(24, 2, 3, 12, 10)
    torch.manual_seed(7)
    out["diffaug_bs"] = DiffAugment(xa, policy="bs").numpy()
    torch.manual_seed(7)
    out["diffaug_u"] = np.stack([torch.rand(2, 1, 1, 1).numpy().reshape(-1), torch.rand(2, 1, 1, 1).numpy().reshape(-1)])
    # PatchSampleF + PatchNCELoss (a14)
    feats = [rand_input(25, 2, 6, 8, 8), rand_input(26, 2, 12, 4, 4)]
    ids = [np.random.RandomState(1).permutation(64)[:16], np.random.RandomState(2).permutation(16)[:16]]
    with quiet():
        Fn = PatchSampleF(use_mlp=False)
        fo, _ = Fn(feats, 16, ids)
    out["psf_ids0"], out["psf_ids1"] = ids
    out["psf_out0"], out["psf_out1"] = fo[0].numpy(), fo[1].numpy()
    torch.manual_seed(9)
    with quiet():
        Fm = PatchSampleF(use_mlp=True, init_type="xavier", init_gain=1.0, nc=32)
        fm, _ = Fm(feats, 16, ids)
    out.update(sd_np(Fm.state_dict(), "psf_mlp."))
    out["psf_mlp_out0"], out["psf_mlp_out1"] = fm[0].detach().numpy(), fm[1].detach().numpy()
    for nm, allneg in (("nce_same", False), ("nce_all", True)):
        o = argparse.Namespace(nce_includes_all_negatives_from_minibatch=allneg, batch_size=2, nce_T=0.07)
        q = fm[0].detach().clone().requires_grad_(True)
        k = torch.roll(fm[0].detach(), 1, 0) * 0.5 + fm[0].detach() * 0.5
        k = k / k.norm(dim=1, keepdim=True)
        loss = PatchNCELoss(o)(q, k)
        loss.mean().backward()
        out[nm + "_q"], out[nm + "_k"] = q.detach().numpy(), k.numpy()
        out[nm + "_loss"], out[nm + "_dq"] = loss.detach().numpy(), q.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **out)
    print("ops.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


def make_step(tag, S, NT, NF, extra):
    """Full SinSKITGModel.optimize_parameters through the reference's own option parser."""
    load_reference()
    cwd = os.getcwd()
    os.chdir(os.environ.get("VTS_REFERENCE_ROOT", "/root/reference"))
    try:
        with quiet():
            from options.train_options import TrainOptions
            from models import create_model
        sys.path.insert(0, ROOT)
        from oracle import skit_oracle as O
        cmd = ("--model sinskitG --gpu_ids -1 --name golden --checkpoints_dir /tmp/vts_golden "
               "--crop_size %d --center_w %d --center_h %d --lambda_G1_lpips 0 --lambda_G2_lpips 0 "
               "--use_vision_aided_loss False --batch_size_G2 %d --add_fake_T_sample_size %d " % (S, S, S, NT, NF)
               + extra).split()
        torch.manual_seed(100)
        with quiet():
            to = TrainOptions()
            to.cmd_line = cmd
            opt = to.parse()
            model = create_model(opt)
            model.setup(opt)
        model.train()
        out = {}
        for nm in ("G", "D", "D2"):
            out.update(sd_np(getattr(model, "net" + nm).state_dict(), nm + "_before."))
        batch = O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True)
        torch.manual_seed(200)
        random.seed(300)
        with quiet():
            model.set_input(batch, phase="train")
            model.optimize_parameters(1)
        # replay the RNG streams the step consumed (DiffAugment: real b,s then fake b,s; random.sample)
        torch.manual_seed(200)
        u = [torch.rand(1, 1, 1, 1).item() for _ in range(4)]
        out["rand_u"] = np.array(u, dtype=np.float32)
        out["fake_ox"] = model.fake_sample_offset_x.numpy().reshape(-1).astype(np.int32)
        out["fake_oy"] = model.fake_sample_offset_y.numpy().reshape(-1).astype(np.int32)
        losses = model.get_current_losses()
        for k, v in losses.items():
            out["loss_" + k] = np.float64(v)
        out["fake_I"], out["fake_I_norm"] = sub(model.fake_I, 2)
        out["fake_T"], out["fake_T_norm"] = sub(model.fake_T, 2)
        out["fake_N"], out["fake_N_norm"] = sub(model.fake_N, 2)
        out["aug_fake_I"], out["aug_fake_I_norm"] = sub(model.aug_fake_I, 2)
        out["pred_fake_T_full"] = model.pred_fake_T_full.numpy()
        for nm in ("G", "D", "D2"):
            net = getattr(model, "net" + nm)
            for k, p in net.named_parameters():
                if p.grad is not None:
                    out["%s_grad.%s" % (nm, k)] = p.grad.numpy()
            # post-Adam weights: flat strided subsample + norm (full copies would triple the file)
            for k, v in net.state_dict().items():
                if "running" in k or "num_batches" in k:
                    out["%s_after.%s" % (nm, k)] = v.numpy()
                else:
                    out["%s_after_sub.%s" % (nm, k)] = v.detach().reshape(-1)[::7].numpy().copy()
                    out["%s_after_norm.%s" % (nm, k)] = np.float64(v.detach().double().norm().item())
        out["meta"] = np.array([S, NT, NF])
        np.savez_compressed(os.path.join(OUT, "step_%s.npz" % tag), **out)
        print("step_%s.npz" % tag, sum(v.nbytes for v in out.values()) / 1e6, "MB raw", {k: round(v, 5) for k, v in losses.items()})
    finally:
        os.chdir(cwd)


def make_metrics():
    """The evaluation metrics that are the reference's OWN code: the surface-normal angle error (models/normal_losses.py on
    model_utils.compute_normal, as compute_evaluation_metric calls it, model_utils.py:531-536) and the clamped touch MSE
    (:520,553-555), on seeded touch patches.  PSNR / SSIM are torchmetrics functions (absent): no fixture."""
    load_reference()
    with quiet():
        from models.model_utils import compute_normal
        from models.normal_losses import compute_surface_normal_angle_error
    g = torch.Generator().manual_seed(21)
    real_T = torch.rand(6, 2, 32, 32, generator=g)
    fake_T = torch.rand(6, 2, 32, 32, generator=g) * 1.4 - 0.2          # leaves [0, 1]: exercises the clamp
    fc = torch.clamp(fake_T, 0, 1)
    ae = compute_surface_normal_angle_error(compute_normal(real_T, scale_nz=1), compute_normal(fc, scale_nz=1), mode="evaluate").mean()
    mse = torch.mean((real_T - fc) ** 2)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), real_T=real_T.numpy(), fake_T=fake_T.numpy(), T_AE=np.float64(ae.item()),
                        T_MSE=np.float64(mse.item()))
    print("metrics.npz: T_AE %.6f deg, T_MSE %.6f" % (ae.item(), mse.item()))


def make_ganloss():
    """GANLoss (models/networks.py:448-542) in every mode it defines, with and without label smoothing (sinskitG_model.py:485-490), on a
    multiscale prediction list and on a bare tensor, real and fake targets: loss values and d(loss.mean())/d(prediction)."""
    N = ref_networks()
    g = torch.Generator().manual_seed(33)
    preds = [torch.randn(3, 1, 9, 9, generator=g) * 1.5, torch.randn(3, 1, 5, 5, generator=g) * 1.5, torch.randn(3, 1, 4, 4, generator=g) * 1.5]
    out = {"pred%d" % i: p.numpy() for i, p in enumerate(preds)}
    for mode in ("nonsaturating", "hinge", "wgan", "wgangp", "lsgan", "vanilla"):
        for smooth in (False, True):
            crit = N.GANLoss(mode, target_real_label=0.8, target_fake_label=0.0) if smooth else N.GANLoss(mode)
            for is_real in (True, False):
                ps = [p.clone().requires_grad_(True) for p in preds]
                loss = crit([[p] for p in ps], is_real)
                loss.mean().backward()
                key = "%s/%d/%d" % (mode, int(smooth), int(is_real))
                out[key + "/multi"] = loss.detach().numpy()
                for i, p in enumerate(ps):
                    out[key + "/grad%d" % i] = p.grad.numpy()
                out[key + "/bare"] = crit([preds[0]], is_real).detach().numpy()      # a list holding one tensor: input[-1]
    np.savez_compressed(os.path.join(OUT, "ganloss.npz"), **out)
    print("ganloss.npz: %d arrays" % len(out))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["networks", "ops", "step", "stylegan2", "options", "step_lpips", "metrics", "ganloss"]
    if "ganloss" in which:
        make_ganloss()
    if "metrics" in which:
        make_metrics()
    if "options" in which:
        make_options()
    if "step_lpips" in which:
        make_step_lpips()
    if "stylegan2" in which:
        make_stylegan2()
    if "networks" in which:
        make_networks()
    if "ops" in which:
        make_ops()
    if "step" in which:
        make_step("resnet", 64, 8, 4, "--netG resnet_9blocks --ngf 8")
        make_step("unet", 256, 8, 4, "--ngf 4")
