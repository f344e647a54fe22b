"""compute_evaluation_metric on the device (visual-tactile-synthesis_b200/eval_metrics.py, csrc/metrics.cu) against the oracle's
restatement of models/model_utils.py:431-561 and against the reference-generated fixture for the metrics that are the reference's
own code (T_AE, T_MSE).  Tolerance 1e-4 relative on every value (fp32 inputs, fp64 accumulation)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(96, 130), (512, 512)])
def test_metrics_match_oracle(golden_dir, hw):
    import vts_b200
    from oracle import skit_oracle as O
    g = torch.Generator().manual_seed(hw[0])
    real_I = torch.rand(1, 3, *hw, generator=g) * 1.8 - 0.9
    fake_I = real_I + 0.2 * torch.randn(real_I.shape, generator=g)          # leaves the real range: exercises the clamp
    z = np.load(os.path.join(golden_dir, "metrics.npz"))
    rT, fT = torch.from_numpy(z["real_T"]), torch.from_numpy(z["fake_T"])
    want = O.evaluation_metrics(real_I, fake_I, rT, fT)
    got = vts_b200.eval_metrics.compute_evaluation_metric(["G"], real_I.cuda(), fake_I.cuda(), rT.cuda(), fT.cuda(),
                                                          eval_metrics=["I_PSNR", "I_SSIM", "T_AE", "T_MSE"], prefix="test_")
    assert set(got) == {"metric_test_I_PSNR", "metric_test_I_SSIM", "metric_test_T_AE", "metric_test_T_MSE"}
    for k, v in want.items():
        assert abs(float(got["metric_test_" + k]) - v) <= 1e-4 * max(1.0, abs(v)), (k, got["metric_test_" + k], v)
    # the reference's own numbers
    assert abs(float(got["metric_test_T_AE"]) - float(z["T_AE"])) < 1e-3 and abs(float(got["metric_test_T_MSE"]) - float(z["T_MSE"])) < 1e-6
    # identical images: SSIM 1, PSNR infinite
    same = vts_b200.eval_metrics.compute_evaluation_metric(["G"], real_I.cuda(), real_I.cuda(), eval_metrics=["I_SSIM"])
    assert abs(float(same["metric_I_SSIM"]) - 1.0) < 1e-5


def test_metrics_lpips_hook_and_unbuilt_metrics_raise():
    import vts_b200
    from oracle import skit_oracle as O
    g = torch.Generator().manual_seed(3)
    a, b = (torch.rand(1, 3, 64, 64, generator=g) * 2 - 1).cuda(), (torch.rand(1, 3, 64, 64, generator=g) * 2 - 1).cuda()
    with pytest.raises(NotImplementedError, match="InceptionV3"):
        vts_b200.eval_metrics.compute_evaluation_metric(["G"], a, b, eval_metrics=["I_SIFID"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vts_b200.eval_metrics.compute_evaluation_metric(["G"], a.cpu(), b.cpu(), eval_metrics=["I_PSNR"])
    crit = vts_b200.lpips_vgg.LPIPS(net="vgg").cuda()
    sdL = O.lpips_random_state(5)
    crit.load_state_dict(sdL, strict=False)
    crit.refresh_packs_once()
    got = vts_b200.eval_metrics.compute_evaluation_metric(["G"], a, b, eval_metrics=["I_LPIPS"], eval_LPIPS=crit)
    want = O.lpips_vgg(sdL, a.cpu(), b.cpu()).mean().item()
    assert abs(float(got["metric_I_LPIPS"]) - want) <= 1e-3 * max(1.0, abs(want))
