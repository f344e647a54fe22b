"""Evaluation metrics of the skitG / sinskitG models on the device:
`compute_evaluation_metric` with the reference's signature and result keys (models/model_utils.py:431-561), as single-pass
reduction kernels (csrc/metrics.cu) over the full-resolution outputs instead of a dozen ATen passes plus torchmetrics.

Built: I_PSNR, I_SSIM (after the reference's min-max rescale of both images to the REAL image's range and clamp of the fake one,
:483-487), T_AE (surface-normal angle error in degrees, models/normal_losses.py:10-33 on compute_normal(., scale_nz=1)), T_MSE, and
I_LPIPS / T_LPIPS through whatever criterion is passed as `eval_LPIPS` (the reference builds lpips.LPIPS(net='alex') for this,
sinskitG_model.py:501; `lpips_vgg.LPIPS` is the one built here).  Not built: I_SIFID / T_SIFID / T_FID — InceptionV3 features plus
a matrix square root (models/sifid.py:102-233, models/tactile_patch_fid.py), whose pretrained network is third-party and not
available offline: asking for them raises.

PSNR and SSIM are torchmetrics functions in the reference (`requirements.txt:18`, unpinned, not installed here): restated from
torchmetrics 0.11 `functional/image/{psnr,ssim}.py` — PARITY UNPINNED for those two (no fixture can be generated); T_AE and T_MSE
are the reference's own code and are pinned by tests/golden/metrics.npz.
"""
import torch

from . import _lib as L

_p = L.ptr

BUILT = ("I_PSNR", "I_SSIM", "I_LPIPS", "T_LPIPS", "T_AE", "T_MSE")
NOT_BUILT = ("I_SIFID", "T_SIFID", "T_FID")


def _acc(device):
    return torch.zeros(1, dtype=torch.float64, device=device)


def minmax(x):
    """-> device fp32 [2] = (min, max) of x."""
    out = torch.tensor([float("inf"), float("-inf")], dtype=torch.float32, device=x.device)
    L.call("skit_metric_minmax", _p(x), x.numel(), _p(out), L.stream())
    return out


def psnr(real, fake, mm=None, data_range=1.0):
    """torchmetrics PSNR(preds, target, data_range): 10 log10(data_range^2 / mse) over all elements; inputs rescaled with `mm`
    (and the fake one clamped) inside the kernel."""
    acc = _acc(real.device)
    L.call("skit_metric_sq_err", _p(real), _p(fake), real.numel(), _p(mm), 0, _p(acc), L.stream())
    mse = acc / real.numel()
    return (10.0 * torch.log10(data_range ** 2 / mse)).float().reshape(())


def ssim(real, fake, mm=None, data_range=1.0):
    """torchmetrics SSIM(preds, target, data_range): Gaussian 11x11 / sigma 1.5 windows, mean of the index map without its
    5-pixel border, averaged over the batch."""
    n, c, h, w = real.shape
    acc = _acc(real.device)
    L.call("skit_metric_ssim", _p(real), _p(fake), n * c, h, w, _p(mm), float(data_range), _p(acc), L.stream())
    return (acc / (n * c * (h - 10) * (w - 10))).float().reshape(())


def normal_angle_error(real_T, fake_T, scale_nz=1.0, clamp_fake=False):
    """mean over patches and pixels of acos(<n_real, n_fake>) in degrees."""
    n, c, h, w = real_T.shape
    assert c == 2
    acc = _acc(real_T.device)
    L.call("skit_metric_normal_angle", _p(real_T), _p(fake_T), n, h, w, float(scale_nz), int(clamp_fake), _p(acc), L.stream())
    return (acc / (n * h * w)).float().reshape(())


def mse_clamped(real, fake):
    acc = _acc(real.device)
    L.call("skit_metric_sq_err", _p(real), _p(fake), real.numel(), None, 1, _p(acc), L.stream())
    return (acc / real.numel()).float().reshape(())


def compute_evaluation_metric(model_names, real_I, fake_I, real_T_concat=None, fake_T_concat=None, eval_metrics=(), eval_LPIPS=None,
                              opt=None, device=None, verbose=False, timing=False, prefix=""):
    """models/model_utils.py:431-561.  Returns {"metric_<prefix><name>": 0-d numpy value}, one device -> host read at the end."""
    bad = [m for m in eval_metrics if m in NOT_BUILT]
    if bad:
        raise NotImplementedError("evaluation metrics %s need the pretrained InceptionV3 network (models/sifid.py), which is a "
                                  "third-party checkpoint outside the B200 path" % bad)
    for t in (real_I, fake_I):
        if not t.is_cuda:
            raise RuntimeError("compute_evaluation_metric (B200 path) needs CUDA tensors; there is no CPU fallback")
    real_I, fake_I = real_I.contiguous().float(), fake_I.contiguous().float()
    out = {}
    if "I_LPIPS" in eval_metrics:
        if eval_LPIPS is None:
            raise ValueError("I_LPIPS needs an LPIPS criterion (eval_LPIPS)")
        out["I_LPIPS"] = eval_LPIPS(real_I, fake_I).mean().reshape(())
    mm = minmax(real_I) if ("I_PSNR" in eval_metrics or "I_SSIM" in eval_metrics) else None
    if "I_PSNR" in eval_metrics:
        out["I_PSNR"] = psnr(real_I, fake_I, mm)
    if "I_SSIM" in eval_metrics:
        out["I_SSIM"] = ssim(real_I, fake_I, mm)
    if real_T_concat is not None:
        rT, fT = real_T_concat.contiguous().float(), fake_T_concat.contiguous().float()
        if "T_LPIPS" in eval_metrics:
            if eval_LPIPS is None:
                raise ValueError("T_LPIPS needs an LPIPS criterion (eval_LPIPS)")
            # compute_touch_lpips_loss on the 224 x 224 nearest-neighbour resize of the clamped patches (:524-529): gx and gy as
            # single-channel images, summed over the patches
            f = torch.nn.functional.interpolate
            r224, f224 = f(rT, (224, 224)), f(fT.clamp(0, 1), (224, 224))
            tot = 0
            for ch in (0, 1):
                tot = tot + eval_LPIPS(f224[:, ch:ch + 1].contiguous(), r224[:, ch:ch + 1].contiguous()).reshape(-1).sum()
            out["T_LPIPS"] = tot.reshape(())
        if "T_AE" in eval_metrics:
            out["T_AE"] = normal_angle_error(rT, fT, scale_nz=1.0, clamp_fake=True)
        if "T_MSE" in eval_metrics:
            out["T_MSE"] = mse_clamped(rT, fT)
    if not out:
        return {}
    vals = torch.stack([v.float() for v in out.values()]).cpu().numpy()      # the only synchronisation
    return {"metric_%s%s" % (prefix, k): vals[i] for i, k in enumerate(out)}
