"""The BaseModel contract the reference's train.py / test.py drive (models/base_model.py:71-230), exercised the way
train.py:35-104 and test.py:65-92 call it, on synthetic batches; and SKITGModel (U-Net generator + per-material style code)
through the train step against the oracle, including a change of style code between replays of the captured step graph."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GATE = 1e-3


def rel(a, b):
    b = b.detach().cpu() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b))
    a, b = a.detach().double().cpu(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_train_loop_body_and_test_loop_of_the_reference(tmp_path):
    import vts_b200
    from oracle import skit_oracle as O
    S, NT, NF = 64, 8, 4
    torch.manual_seed(0)
    opt = vts_b200.default_options(batch_size_G2=NT, add_fake_T_sample_size=NF, checkpoints_dir=str(tmp_path), name="exp",
                                   n_epochs=1, n_epochs_decay=3, cuda_graph_warmup=1)
    model = vts_b200.SinSKITGModel(opt)
    dataset = [O.synthetic_batch(S, NT=NT, seed=s, ellipse_mask=True) for s in range(3)]
    total_iters = 0
    for epoch in range(opt.epoch_count, opt.n_epochs + opt.n_epochs_decay + 1):
        # ---- train.py:35-83 (train_model)
        model.train()
        for i, data in enumerate(dataset):
            total_iters += data["S"].size(0)
            if epoch == opt.epoch_count and i == 0:
                model.setup(opt)
                model.parallelize()
            model.set_input(data, phase="train", verbose=False)
            model.optimize_parameters(epoch)
            torch.cuda.synchronize()
            losses = model.get_current_losses()
            assert list(losses)[:4] == ["l_G_GAN", "l_D_real_I", "l_D_fake_I", "l_D_I_grad_penalty"]
            assert {"l_G_L1", "l_G2_GAN", "l_D_real_T_concat", "l_D_fake_T_concat", "l_D_T_grad_penalty", "l_D_more_fake_T", "l_G2_L1"} <= set(losses)
            assert all(np.isfinite(v) for v in losses.values())
            vis = model.get_current_visuals()
            assert {"real_S", "M", "real_I", "fake_I", "fake_gx", "fake_gy", "fake_N", "pred_fake_I", "aug_fake_I", "aug_real_I"} <= set(vis)
            assert vis["fake_I"].shape == (1, 3, S, S) and vis["fake_gx"].shape == (1, 1, S, S) and vis["pred_fake_I"].shape[1] == 1
            model.save_networks("latest")
        # ---- train.py:85-101 (validation pass) and :156-205
        model.eval()
        model.set_input(dataset[0], phase="val", verbose=False)
        model.test()
        assert model.get_image_paths() == ["syn.png"]
        if epoch == opt.epoch_count:
            assert model.get_current_metrics() == {}
        res = model.compute_current_metrics()          # what the reference computes inside get_current_visuals
        assert set(res) == {"metric_I_PSNR", "metric_I_SSIM", "metric_T_AE", "metric_T_MSE"}     # validation phase: no 'train_' prefix
        mets = model.get_current_metrics()
        assert {"m_I_PSNR", "m_I_SSIM", "m_T_AE", "m_T_MSE"} <= set(mets) and all(np.isfinite(v) for v in mets.values())
        assert 0.0 < mets["m_I_SSIM"] < 1.0 and 0.0 <= mets["m_T_AE"] <= 180.0
        model.save_networks(epoch)
        model.update_learning_rate()
        # LambdaLR after k steps: lambda_rule(k) (networks.py:161-165)
        k = epoch - opt.epoch_count + 1
        assert abs(model.lr_factor - (1.0 - max(0, k + opt.epoch_count - opt.n_epochs) / float(opt.n_epochs_decay + 1))) < 1e-12
    assert model._graph is not None      # the loop ran on the captured step
    # ---- test.py:65-92: a test-time model loads G from the checkpoint directory in setup() and reproduces the forward
    topt = vts_b200.default_options(isTrain=False, checkpoints_dir=str(tmp_path), name="exp", epoch="latest")
    tm = vts_b200.SinSKITGModel(topt)
    tm.setup(topt)
    tm.parallelize()
    tm.eval()
    model.set_input(dataset[1], phase="val")
    ref = [t.clone() for t in model.test(timing=False)]
    tm.set_input({k: dataset[1][k] for k in ("S", "M", "name", "S_paths", "augmentation_params")}, phase="test")
    out = tm.test()
    torch.cuda.synchronize()
    for a, b in zip(out, ref):
        assert rel(a, b) < 1e-6
    sd = torch.load(str(tmp_path / "exp" / "latest_net_G.pth"))
    assert "model.1.weight" in sd and "model.30.bias" in sd     # reference key names: released checkpoints interchange


def test_guards_mirror_the_reference_or_fail_loudly():
    import vts_b200
    from oracle import skit_oracle as O
    with pytest.raises(ValueError, match="lpips_state"):
        vts_b200.SinSKITGModel(vts_b200.reference_default_options())
    with pytest.raises(NotImplementedError, match="DiffAugment"):
        vts_b200.SinSKITGModel(vts_b200.default_options(diffaugment="bst"))
    m = vts_b200.SinSKITGModel(vts_b200.default_options(batch_size_G2=4, add_fake_T_sample_size=2))
    b = O.synthetic_batch(64, NT=4, seed=0)
    bad = dict(b)
    c = b["T_coords"].clone()
    c[0, :, 5] = 2.0                      # resize_ratio 2 -> cutout 16 != patch size 32 (model_utils.py:337-341 resizes)
    bad["T_coords"] = c
    with pytest.raises(NotImplementedError, match="cutout"):
        m.set_input(bad)
    # use_bg_mask off: the mask multiplies are skipped, i.e. the result equals an all-ones mask
    torch.manual_seed(1)
    m0 = vts_b200.SinSKITGModel(vts_b200.default_options(isTrain=False, use_bg_mask=False))
    be = O.synthetic_batch(64, NT=4, seed=0, ellipse_mask=True)
    m0.set_input({"S": be["S"]}, phase="test")
    a = m0.test()[0].clone()
    m0.opt.use_bg_mask = True
    m0.set_input({"S": be["S"], "M": torch.ones_like(be["M"])}, phase="test")
    assert rel(m0.test()[0], a) < 1e-6


@pytest.mark.parametrize("graph", [False, True])
def test_skitg_model_style_code_train_step_vs_oracle(graph):
    """skitG (models/skitG_model.py): unet256_custom with use_style_code (concat / tile, 512-d, innermost level), fp16 code
    promoted to fp32 like torch.cat does.  Step 1 vs the oracle; with the CUDA graph on, the material (style code) changes
    between replays and the generator output must follow it."""
    import vts_b200
    from oracle import skit_oracle as O
    S, NT, NF = 256, 8, 4
    torch.manual_seed(3)
    opt = vts_b200.default_options(model="skitG", netG="unet256_custom", ngf=10, ndf=8, use_style_code=True, batch_size_G2=NT,
                                   add_fake_T_sample_size=NF, cuda_graph=graph, cuda_graph_warmup=1)
    m = vts_b200.SKITGModel(opt)
    assert "up7.model.1.weight" in m.netG.state_dict() and m.netG.state_dict()["up7.model.1.weight"].shape[0] == 592
    sds = [{k: v.detach().cpu().clone() for k, v in net.state_dict().items()} for net in (m.netG, m.netD, m.netD2)]
    g = torch.Generator().manual_seed(5)
    codes = [torch.randn(1, 512, generator=g).half() for _ in range(3)]
    batch = O.synthetic_batch(S, NT=NT, seed=0, ellipse_mask=True)
    batch["style_code"] = codes[0]
    rs = np.random.RandomState(2)
    rand = dict(real_b=[0.3], real_s=[0.8], fake_b=[0.6], fake_s=[0.2],
                fake_ox=rs.randint(0, S - 32, NF).astype(np.int32), fake_oy=rs.randint(0, S - 32, NF).astype(np.int32))
    m.set_input(batch)
    assert m.M_T is m.M
    m.optimize_parameters(1, rand=rand)
    torch.cuda.synchronize()
    cfg = O.StepConfig(netG="unet256_custom", batch_size_G2=NT, add_fake_T_sample_size=NF)
    sdG, sdD, sdD2 = [copy.deepcopy(s) for s in sds]
    res = O.train_step(cfg, sdG, sdD, sdD2, {}, O.step_inputs_from_batch(batch), rand, step=1)
    losses = m.current_losses()
    for k, v in res["losses"].items():
        assert abs(losses[k] - v) <= GATE * max(1.0, abs(v)), (k, losses[k], v)
    assert rel(m.fake_I, res["fake_I"]) < GATE and rel(m.fake_T, res["fake_T"]) < GATE
    for k, p in m.netG.named_parameters():
        if k in res["grads_G"]:
            g_ref, wk = res["grads_G"][k], k.replace(".bias", ".weight")
            if k.endswith(".bias") and wk in res["grads_G"] and g_ref.norm() < 1e-3 * res["grads_G"][wk].norm():
                continue    # a conv bias feeding an InstanceNorm: mathematically zero, the reference holds rounding noise
            assert rel(p.grad, g_ref) < 3e-2, k
        else:
            assert "style_code_mapping" in k and float(p.grad.abs().max()) == 0.0, k    # never used in 'tile' mode (A.1)
    # the next materials: the step (eager or captured-and-replayed) must read the CURRENT style code
    outs = []
    for i in (1, 2, 1):
        batch["style_code"] = codes[i]
        m.set_input(batch)
        m.optimize_parameters(1, rand=rand)
        torch.cuda.synchronize()
        sd_now = {k: v.detach().cpu().clone() for k, v in m.netG.state_dict().items()}
        outs.append((m.fake_I.clone(), sd_now))
    if graph:
        assert m._graph is not None
    # compare the last generator output with the oracle forward under the weights it was produced with: the weights the
    # step started from are those saved after the previous step
    x = torch.cat([batch["S"] * batch["M"], O.spe_grid(S, S, 4, 1)], 1)
    ref = O.unet_custom_forward(outs[1][1], x, style_code=codes[1].float())[:, 0:3] * batch["M"]
    assert rel(outs[2][0], ref) < GATE
    wrong = O.unet_custom_forward(outs[1][1], x, style_code=codes[2].float())[:, 0:3] * batch["M"]
    assert rel(outs[2][0], wrong) > 10 * rel(outs[2][0], ref)
