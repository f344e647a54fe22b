"""CPU: host-side logic of the drop-in boundary — factories, state_dict compatibility with the real
reference (golden fixtures), error behaviour, coordinate arithmetic, flat parameter buckets, LR rule.
The CUDA path itself is covered by the -m gpu tests; here we also check it fails LOUDLY on CPU."""
import argparse
import os
import random

import numpy as np
import pytest
import torch

import vts_b200
from vts_b200 import networks as N
from vts_b200 import model_utils as MU
from oracle import skit_oracle as O


def z(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def ns(**kw):
    return argparse.Namespace(gan_mode="nonsaturating", **kw)


def test_state_dict_keys_and_shapes_match_reference(golden_dir):
    g = z(golden_dir, "networks.npz")
    for prefix, net in (
        ("Gres.", N.define_G(9, 5, 8, "resnet_9blocks", "instance", False, "xavier", 0.02, False, False, [], ns())),
        ("D_before.", N.define_D(7, 8, "multiscale", 3, "batch", "xavier", 0.02, False, 3, [], ns())),
        ("Dbasic.", N.define_D(4, 8, "basic", 3, "batch", "xavier", 0.02, False, 3, [], None)),
    ):
        ref = {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}
        sd = net.state_dict()
        assert set(sd) == set(ref), prefix
        for k, v in sd.items():
            assert tuple(v.shape) == ref[k].shape, (prefix, k)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in ref.items()})  # released checkpoints stay loadable


def test_parameter_counts_match_survey():
    G = N.define_G(9, 5, 64, "resnet_9blocks", "instance", False, "xavier", 0.02, False, False, [], ns())
    assert abs(sum(p.numel() for p in G.parameters()) / 1e6 - 11.4) < 0.05  # 11.4 M
    D = N.define_D(4, 64, "multiscale", 3, "batch", "xavier", 0.02, False, 3, [], ns())
    assert abs(sum(p.numel() for p in D.parameters()) / 1e6 - 3 * 2.77) < 0.05  # 3 x 2.77 M


def test_factory_error_behaviour_mirrors_reference():
    with pytest.raises(NotImplementedError, match=r"Generator model name \[nope\] is not recognized"):
        N.define_G(9, 5, 64, "nope", "instance", False, "xavier", 0.02, False, False, [], ns())
    with pytest.raises(NotImplementedError, match=r"Discriminator model name \[nope\] is not recognized"):
        N.define_D(4, 64, "nope", 3, "batch", "xavier", 0.02, False, 3, [], ns())
    with pytest.raises(NotImplementedError, match=r"projection model name \[nope\] is not recognized"):
        N.define_F(3, "nope", opt=ns(netF_nc=256))
    with pytest.raises(NotImplementedError, match="gan mode"):
        N.GANLoss("nope")
    with pytest.raises(NotImplementedError, match="learning rate policy"):
        N.get_scheduler(torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=1), ns(lr_policy="nope"))


def test_xavier_init_gain_and_zero_bias():
    torch.manual_seed(0)
    G = N.define_G(9, 5, 64, "resnet_9blocks", "instance", False, "xavier", 0.02, False, False, [], ns())
    w = G.state_dict()["model.12.conv_block.1.weight"]
    assert abs(w.std().item() / (0.02 * (2.0 / (2304 * 2)) ** 0.5) - 1) < 0.05     # networks.py:204-222, gain 0.02
    assert float(G.state_dict()["model.12.conv_block.1.bias"].abs().max()) == 0.0
    D = N.define_D(4, 8, "multiscale", 3, "batch", "xavier", 0.02, False, 3, [], ns())
    bnw = D.state_dict()["layer0.3.weight"]
    assert abs(bnw.mean().item() - 1) < 0.05 and bnw.std().item() < 0.06              # N(1, 0.02)


def test_no_cpu_fallback_fails_loudly():
    G = N.define_G(9, 5, 8, "resnet_9blocks", "instance", False, "xavier", 0.02, False, False, [], ns())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G(torch.zeros(1, 9, 32, 32))
    D = N.define_D(4, 8, "multiscale", 3, "batch", "xavier", 0.02, False, 3, [], ns())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D(torch.zeros(1, 4, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        N.GANLoss("nonsaturating")([[torch.zeros(1, 1, 4, 4)]], True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MU.get_patch_in_input(torch.zeros(1, 3, 64, 64), np.zeros((1, 2, 8)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vts_b200.PatchNCELoss(ns(nce_includes_all_negatives_from_minibatch=False, batch_size=1, nce_T=0.07))(torch.zeros(4, 8), torch.zeros(4, 8))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            vts_b200.SinSKITGModel(vts_b200.default_options())


def test_patch_coordinates_match_reference_golden(golden_dir):
    g = z(golden_dir, "ops.npz")
    ox, oy, cs = MU.find_coords_for_patch(g["gather_coords"])
    assert np.array_equal(ox, g["gather_ox"].reshape(-1)) and np.array_equal(oy, g["gather_oy"].reshape(-1))
    assert np.array_equal(cs, g["gather_cs"].reshape(-1))
    # half-to-even rounding of the float64 arithmetic (model_utils.py:37-57)
    c = np.array([[[10.5, 11.5, 64, 64, 32, 1, 0, 0], [2.0, 3.0, 64, 64, 32, 2, 5, 7]]])
    ox, oy, cs = MU.find_coords_for_patch(c)
    assert list(ox) == [10, 4] and list(oy) == [12, 6] and list(cs) == [32, 16]
    # random-mode candidate table: same order and same picks as the reference for the same `random` state
    M = torch.from_numpy(g["rand_M"])
    random.seed(5)
    rox, roy = MU.random_patch_offset_table(M).sample(12)
    assert np.array_equal(rox, g["rand_ox"]) and np.array_equal(roy, g["rand_oy"])
    ones = torch.ones(1, 1, 40, 52)
    t = MU.random_patch_offset_table(ones)
    assert len(t) == (40 - 14) * (52 - 14)
    random.seed(1)
    a = t.sample(5)
    random.seed(1)
    b = O.random_patch_offsets(ones, 5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # the bit-map form of the table (what the device kernel returns and the device dataset stores with each item): same picks as
    # the reference golden when built from the host table's own box
    host = MU.random_patch_offset_table(M)
    oh, ow = host.box.shape
    words = (ow + 31) // 32
    padded = np.zeros((oh, words * 32), np.uint8)
    padded[:, :ow] = host.box
    bits = np.packbits(padded, axis=1, bitorder="little")
    tb = MU.offset_table_from_arrays(bits, host.box.sum(1).astype(np.int32), M.shape[-2], M.shape[-1])
    random.seed(5)
    box, boy = tb.sample(12)
    assert np.array_equal(box, g["rand_ox"]) and np.array_equal(boy, g["rand_oy"])
    assert len(tb) == len(host) and np.array_equal(tb.rows, host.rows) and np.array_equal(tb.cols, host.cols)


def test_positional_encoding_matches_reference_golden(golden_dir):
    g = z(golden_dir, "ops.npz")
    np.testing.assert_allclose(MU.spe_grid(20, 28, 4, 2).numpy(), g["spe_20x28"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(MU.spe_grid(1100, 8, 4, 1).numpy(), g["spe_1100"], rtol=1e-5, atol=1e-5)


def test_flat_parameter_bucket_views():
    D = N.define_D(4, 8, "multiscale", 3, "batch", "xavier", 0.02, False, 3, [], ns())
    before = {k: v.clone() for k, v in D.state_dict().items()}
    D.flatten_parameters()
    total = sum(p.numel() for p in D.parameters())
    assert D.flat_param.numel() == total == D.flat_grad.numel()
    for k, v in D.state_dict().items():
        assert torch.equal(v, before[k]), k
    p = next(D.parameters())
    D.flat_param.add_(1.0)                       # the optimiser kernel writes the bucket: parameters must see it
    assert torch.equal(p.data, before["layer0.0.weight"] + 1.0)
    D.flat_grad.fill_(2.0)
    assert float(p.grad.mean()) == 2.0
    D.zero_grad()
    assert float(D.flat_grad.abs().max()) == 0.0 and p.grad is not None
    D.load_state_dict(before)                    # load copies in place: still views of the bucket
    assert p.data.untyped_storage().data_ptr() == D.flat_param.untyped_storage().data_ptr()
    D.ensure_flat()
    assert D.flat_param.numel() == total


def test_linear_lr_rule_and_default_options():
    opt = vts_b200.default_options()
    assert (opt.beta1, opt.beta2, opt.lr, opt.lr_G2) == (0.0, 0.99, 1e-3, 5e-4)     # sinskitG_model.py:331-332,596-598
    assert (opt.lambda_G1_L1, opt.lambda_G2_GAN, opt.lambda_G2_L1) == (100.0, 5.0, 10.0)
    for e in (0, 4, 5, 100, 404):
        want = O.linear_lr_factor(e, opt.epoch_count, opt.n_epochs, opt.n_epochs_decay)
        got = 1.0 - max(0, e + opt.epoch_count - opt.n_epochs) / float(opt.n_epochs_decay + 1)
        assert abs(want - got) < 1e-12
    sched = N.get_scheduler(torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=1.0), ns(lr_policy="linear", epoch_count=1, n_epochs=5, n_epochs_decay=400))
    assert sched.get_last_lr()[0] == 1.0


def test_tappable_layers_cover_cut_defaults():
    G = N.define_G(9, 5, 8, "resnet_9blocks", "instance", False, "xavier", 0.02, False, False, [], ns())
    assert {0, 4, 8, 12, 16} <= G.tappable_layers()


def _sg2_opt(netG, res):
    return argparse.Namespace(load_size=res, crop_size=res, stylegan2_G_num_downsampling=1, netG=netG)


@pytest.mark.parametrize("tag,netG,res", [("sg2", "stylegan2", 128), ("sg2small", "smallstylegan2", 64)])
def test_stylegan2_state_dict_matches_reference(golden_dir, tag, netG, res):
    """define_G('stylegan2' | 'smallstylegan2'): parameter / buffer names and shapes of the real reference (fixture written by
    oracle/make_golden.py), the blur buffers' values, the channel table, and the loud CPU failure."""
    g = z(golden_dir, "stylegan2.npz")
    ref = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + ".")}
    G = N.define_G(9, 5, 4, netG, "instance", False, "xavier", 0.02, False, False, [], _sg2_opt(netG, res))
    sd = G.state_dict()
    assert sorted(sd.keys()) == sorted(ref.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref[k].shape), k
        if k.endswith("kernel"):
            np.testing.assert_allclose(v.numpy(), ref[k], rtol=0, atol=1e-7)
    G.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in ref.items()})
    assert vts_b200.sg2_generator.channel_table(64)[512] == 64 and vts_b200.sg2_generator.channel_table(64)[256] == 128
    assert vts_b200.sg2_generator.channel_table(10)[1024] == 5      # int(round(16 * 10 / 32))
    with pytest.raises(KeyError):                                    # crop 1536 -> 2048: no such resolution, like the reference (:820)
        N.define_G(9, 5, 64, "stylegan2", "instance", False, "xavier", 0.02, False, False, [], _sg2_opt("stylegan2", 1536))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G(torch.zeros(1, 9, res, res))


def test_lpips_module_keys_and_cpu_behaviour():
    """vts_b200.lpips_vgg.LPIPS carries the lpips package's state_dict keys (so its checkpoints load), is frozen, and fails
    loudly without a CUDA device; the oracle's random state uses the same keys."""
    m = vts_b200.lpips_vgg.LPIPS(net="vgg")
    sd = m.state_dict()
    want = set(O.lpips_random_state(0).keys())
    assert want <= set(sd.keys())
    assert {"lins.%d.model.1.weight" % k for k in range(5)} <= set(sd.keys())
    assert {"scaling_layer.shift", "scaling_layer.scale"} <= set(sd.keys())
    assert sd["net.slice1.0.weight"].shape == (64, 3, 3, 3) and sd["net.slice5.28.weight"].shape == (512, 512, 3, 3)
    assert sd["lin3.model.1.weight"].shape == (1, 512, 1, 1)
    assert not any(p.requires_grad for p in m.parameters()) and not m.training
    np.testing.assert_allclose(sd["scaling_layer.shift"].flatten().numpy(), O.LPIPS_SHIFT, atol=1e-7)
    with pytest.raises(NotImplementedError):
        vts_b200.lpips_vgg.LPIPS(net="alex")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 32, 32), torch.zeros(1, 3, 32, 32))


def test_default_options_against_the_reference_parser(golden_dir):
    """Every option the step reads carries the reference parser's default (fixture: oracle/make_golden.py options), except
    the documented benchmark-architecture / third-party-loss overrides; reference_default_options() restores those."""
    import json
    ref = json.load(open(os.path.join(golden_dir, "options.json")))
    mine = vars(vts_b200.default_options())
    differs = {k for k in mine if k in ref and ref[k] != mine[k]}
    assert differs == {"checkpoints_dir", "gpu_ids", "name", "netG", "ngf", "ndf", "lambda_G1_lpips", "lambda_G2_lpips",
                       "use_vision_aided_loss"}
    r = vars(vts_b200.reference_default_options())
    for k in ("netG", "ngf", "ndf", "lambda_G1_lpips", "lambda_G2_lpips"):
        assert r[k] == ref[k], k
    assert r["use_vision_aided_loss"] is False and ref["use_vision_aided_loss"] is True


def test_first_of_permutation_is_a_uniform_partial_shuffle():
    """PatchSampleF's id draw without the O(H*W) permutation: distinct ids, in range, seedable through np.random.seed like the
    reference's np.random.permutation, uniform marginals (chi-square-ish bound) and uniform first element."""
    np.random.seed(3)
    a = N.first_of_permutation(1000, 256)
    np.random.seed(3)
    b = N.first_of_permutation(1000, 256)
    assert (a == b).all() and len(set(a.tolist())) == 256 and a.min() >= 0 and a.max() < 1000
    assert sorted(N.first_of_permutation(5, 64).tolist()) == [0, 1, 2, 3, 4]          # p > n: a full permutation
    counts, first = np.zeros(20), np.zeros(20)
    for _ in range(4000):
        s = N.first_of_permutation(20, 5)
        counts[s] += 1
        first[s[0]] += 1
    assert np.abs(counts / 4000 - 0.25).max() < 0.04 and np.abs(first / 4000 - 0.05).max() < 0.02
