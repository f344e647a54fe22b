"""SinSKITGModel / SKITGModel on the B200 path: set_input / forward / optimize_parameters / test
with the reference's method names and loss algebra, run as explicit kernel launches (no autograd).

Reference: models/sinskitG_model.py — forward :1293-1344, optimize_parameters :601-700,
compute_D1_loss :1346-1407, compute_D2_loss :1409-1617, compute_G1_loss :1660-1726,
compute_G2_loss :1728-1842, set_input :702-793, optimizers :590-599;  models/skitG_model.py
(:1284-1336 forward, style code / M_T) is the multi-material twin whose own optimize_parameters is
broken as shipped — SKITGModel here reuses the sinskitG step.
The LPIPS-VGG16 terms run on the library's own kernels (lpips_vgg.py) and need the package's checkpoint passed as
`opt.lpips_state` (no pretrained weights exist offline: random weights only behind `opt.allow_random_lpips`); the
vision-aided (CLIP) discriminator and the CLIP style encoder are third-party networks outside the hot path
(SURVEY.md §8f): `use_vision_aided_loss=True` raises, the style code is taken precomputed or from a user-supplied encoder.
The BaseModel contract train.py / test.py drive (models/base_model.py:71-230: setup, parallelize, train / eval,
get_current_visuals / losses / metrics, get_image_paths, update_learning_rate, save / load_networks) is mirrored below;
`integration/b200sinskitG_model.py` is the adapter the reference's `models/__init__.py:54-67` discovers.
"""
import collections
import warnings
import argparse
import math
import os
import random

import numpy as np
import torch

from . import _lib as L_
from . import networks, ops
from .model_utils import find_coords_for_patch, offset_table_from_arrays, random_patch_offset_table, spe_grid


def default_options(**kw):
    """The options the step reads.  Every value is the reference parser's default (models/sinskitG_model.py:50-357,
    options/base_options.py, options/train_options.py; pinned by tests/golden/options.json) EXCEPT the benchmark
    architecture of BASELINE.json configs[1]: netG resnet_9blocks / ngf 64 / ndf 64 (reference: unet256_custom / 10 / 8) and the
    third-party terms off: lambda_G1_lpips 0 / lambda_G2_lpips 0 (reference: 1 / 10) and use_vision_aided_loss False
    (reference: True).  `reference_default_options()` gives the reference's own defaults."""
    o = dict(
        model="sinskitG", isTrain=True, gpu_ids=[0],
        netG="resnet_9blocks", ngf=64, normG="instance", no_dropout=True, no_antialias=False, no_antialias_up=False,
        netD="multiscale", netD2="multiscale", ndf=64, normD="batch", n_layers_D=3, num_D=3,
        init_type="xavier", init_gain=0.02, input_nc=1, output_nc=5, sketch_nc=1, image_nc=3, touch_nc=2,
        num_D_D1=3, num_D_D2=3, n_layers_D2=3, use_cGAN=True, use_cGAN_G2_S=True, use_cGAN_G2_I=True,
        use_positional_encoding=True, use_bg_mask=True, use_diffaug=True, diffaugment="bs", use_more_fakeT=True,
        gan_mode="nonsaturating",
        lambda_G1_GAN=1.0, lambda_G1_L1=100.0, lambda_G1_lpips=0.0, lambda_G2_GAN=5.0, lambda_G2_L1=10.0,
        lambda_G2_lpips=0.0, lambda_G2_GAN_feat=1.0, smooth_GAN_label=True, use_vision_aided_loss=False,
        num_layer_separate=4, use_style_code=False, style_code_mode="concat", style_code_mapping_mode="tile",
        style_code_dim=512, num_layer_style_code=1,
        batch_size=1, batch_size_G2=64, add_fake_T_sample_size=32, scale_nz=0.25, T_resolution_multiplier=1,
        lr=1e-3, lr_G2=5e-4, beta1=0.0, beta2=0.99, lr_policy="linear", n_epochs=5, n_epochs_decay=400, epoch_count=1,
        run_full_res_D2=False,  # the reference's visualisation-only netD2(full image) pass (:1495); off on the hot path
        checkpoints_dir="./checkpoints", name="experiment",
        # PatchNCE (models/patchnce.py + PatchSampleF; dead code in the reference, wired CUT-style here: DESIGN.md section 4.3)
        lambda_NCE=0.0, nce_layers="0,4,8,12,16", num_patches=256, nce_T=0.07, netF="sample", netF_nc=256,
        nce_includes_all_negatives_from_minibatch=False,
        cuda_graph=True,         # replay the whole train step as one CUDA graph once shapes are stable
        cuda_graph_warmup=2,     # eager steps before capture (lazy weight packs, kernel attributes, NCCL warm-up)
        lpips_state=None,        # state_dict of lpips.LPIPS(net='vgg') (the package's own checkpoint keys)
        allow_random_lpips=False,  # benchmarks / parity tests only: run the LPIPS terms on randomly initialised VGG16 weights
        continue_train=False, epoch="latest", verbose=False, pretrained_name=None, train_for_each_epoch=True,
        positional_encoding_mode="spe", positional_encoding_dim=4,
    )
    o.update(kw)
    return argparse.Namespace(**o)


def reference_default_options(**kw):
    """default_options() with the reference's own architecture and loss defaults: the unet256_custom generator (ngf 10),
    PatchGANs with ndf 8, LPIPS-VGG16 terms on (1, 10).  The vision-aided discriminator stays off (not built)."""
    o = dict(netG="unet256_custom", ngf=10, ndf=8, lambda_G1_lpips=1.0, lambda_G2_lpips=10.0)
    o.update(kw)
    return default_options(**o)


def _opt(opt, name, default):
    """Options the reference's parsers do not define (B200-path extras) fall back to their default_options() value."""
    return getattr(opt, name, default)


class SinSKITGModel:
    loss_names = ["D_fake_I", "D_real_I", "D_fake_T_concat", "D_more_fake_T", "D_real_T_concat",
                  "G_GAN", "G_L1", "G2_GAN", "G2_L1"]
    model_names = ["G", "D", "D2"]
    model_name = "sinskitG"

    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        """The B200-path extras on top of the reference model's own options (models/sinskitG_model.py:34-357; the adapter in
        integration/ chains the reference's setter first).  Existing arguments are left alone."""
        have = {a.dest for a in parser._actions}

        def add(name, **kw):
            if name.lstrip("-") not in have:
                parser.add_argument(name, **kw)

        b = lambda v: str(v).lower() in ("1", "true", "yes")   # noqa: E731
        add("--cuda_graph", type=b, default=True, help="replay the train step / test forward as one CUDA graph")
        add("--cuda_graph_warmup", type=int, default=2, help="eager steps before the graph capture")
        add("--run_full_res_D2", type=b, default=False, help="also run the reference's visualisation-only netD2(full image) pass")
        add("--lambda_NCE", type=float, default=0.0, help="weight of the CUT-style PatchNCE term (0 = off, the reference's behaviour)")
        add("--nce_layers", type=str, default="0,4,8,12,16")
        add("--num_patches", type=int, default=256)
        add("--nce_T", type=float, default=0.07)
        add("--netF", type=str, default="sample")
        add("--netF_nc", type=int, default=256)
        add("--nce_includes_all_negatives_from_minibatch", type=b, default=False)
        add("--allow_random_lpips", type=b, default=False)
        return parser

    def __init__(self, opt, dist_ctx=None):
        self.opt = opt
        self.isTrain = opt.isTrain
        self.gpu_ids = opt.gpu_ids
        self.dist = dist_ctx
        if not torch.cuda.is_available():
            raise RuntimeError("SinSKITGModel (B200 path) needs a CUDA device; there is no CPU fallback")
        dev_index = opt.gpu_ids[0] if len(opt.gpu_ids) else 0
        self.device = torch.device("cuda", dev_index)
        torch.cuda.set_device(self.device)
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.image_paths = []
        self.optimizers, self.schedulers = [], []     # Adam lives in the fused kernel; kept for the BaseModel attribute contract
        self.metric = 0
        self.metric_names = []
        if opt.use_vision_aided_loss:
            raise NotImplementedError("the vision-aided discriminator (CLIP / DINO backbones) is a third-party network outside the "
                                      "B200 hot path; run with --use_vision_aided_loss False")
        if opt.T_resolution_multiplier != 1:
            raise NotImplementedError("T_resolution_multiplier != 1 (real bicubic resampling) is not built")
        if opt.batch_size != 1:
            raise NotImplementedError("batch_size is forced to 1 by the reference (sinskitG_model.py:342)")
        if _opt(opt, "use_diffaug", False) and _opt(opt, "diffaugment", "bs") != "bs":
            raise NotImplementedError("DiffAugment policy %r: only 'bs' (brightness + saturation, the model default, "
                                      "sinskitG_model.py:226-231) is built on the B200 path" % opt.diffaugment)
        if _opt(opt, "use_positional_encoding", False) and _opt(opt, "positional_encoding_mode", "spe") != "spe":
            raise NotImplementedError("positional_encoding_mode %r: only 'spe' is built" % opt.positional_encoding_mode)
        # channel counts / discriminator depth under the reference parser's names (sketch_nc, image_nc, touch_nc, num_D_D1,
        # num_D_D2, n_layers_D2: sinskitG_model.py:137-200,505-573), falling back to default_options()'s older spellings
        s_nc = self.sketch_nc = _opt(opt, "sketch_nc", _opt(opt, "input_nc", 1))
        out_nc = _opt(opt, "image_nc", 3) + _opt(opt, "touch_nc", 2) if hasattr(opt, "image_nc") else _opt(opt, "output_nc", 5)
        gan_mode = _opt(opt, "gan_mode", "nonsaturating")
        if gan_mode not in ("nonsaturating", "hinge"):
            # the per-sample modes are the ones the reference's own step survives: with 'lsgan' / 'vanilla' / 'wgan(gp)' GANLoss returns a
            # 0-d tensor and compute_G2_loss fails on len(loss_G2_GAN) (sinskitG_model.py:1782-1784); wgangp also needs a double backward
            raise NotImplementedError("gan_mode %r: the train step is built for 'nonsaturating' (the model default) and 'hinge'; "
                                      "networks.GANLoss itself supports every mode" % gan_mode)
        if (s_nc, out_nc) != (1, 5):
            raise NotImplementedError("the explicit step is built for sketch_nc 1, image_nc 3, touch_nc 2 (the model defaults)")
        if not (_opt(opt, "use_cGAN", True) and _opt(opt, "use_cGAN_G2_S", True) and _opt(opt, "use_cGAN_G2_I", True)):
            raise NotImplementedError("the explicit step is built for the conditional discriminators (use_cGAN / use_cGAN_G2_S / "
                                      "use_cGAN_G2_I = True, the model defaults)")
        g_in = s_nc + (2 * _opt(opt, "positional_encoding_dim", 4) if opt.use_positional_encoding else 0)
        gpu = [dev_index]
        if "stylegan2" in opt.netG:
            raise NotImplementedError("netG=%r: the StyleGAN2 generator (forward + explicit backward, sg2_generator.py) is available "
                                      "through networks.define_G, but its decoder is hard-wired to 3 output channels "
                                      "(stylegan_networks.py:892), which does not feed the 5-channel skitG step" % opt.netG)
        self.netG = networks.define_G(g_in, out_nc, opt.ngf, opt.netG, opt.normG, not opt.no_dropout, opt.init_type,
                                      opt.init_gain, opt.no_antialias, opt.no_antialias_up, gpu, opt,
                                      num_layer_separate=getattr(opt, "num_layer_separate", 4))
        self.netG.flatten_parameters()
        if self.isTrain:
            self.netD = networks.define_D(s_nc + 3, opt.ndf, opt.netD, opt.n_layers_D, opt.normD, opt.init_type,
                                          opt.init_gain, opt.no_antialias, _opt(opt, "num_D_D1", _opt(opt, "num_D", 3)), gpu, opt)
            self.netD2 = networks.define_D(2 + s_nc + 3 + 1, opt.ndf, opt.netD2, _opt(opt, "n_layers_D2", opt.n_layers_D), opt.normD,
                                           opt.init_type, opt.init_gain, opt.no_antialias, _opt(opt, "num_D_D2", _opt(opt, "num_D", 3)), gpu, opt)
            for net in (self.netD, self.netD2):
                if not isinstance(net, networks.MultiscaleDiscriminator):
                    raise NotImplementedError("the explicit train step is built for netD/netD2 = 'multiscale' (the model default)")
                net.flatten_parameters()
            self.step_count = 0
            self.lr_factor = 1.0
            self.lpips = None
            if opt.lambda_G1_lpips > 0 or opt.lambda_G2_lpips > 0:
                # criterionLPIPS_vgg (sinskitG_model.py:495).  The pretrained VGG16 / lin checkpoints are not available offline:
                # random weights unless `opt.lpips_state` (a state_dict with the lpips package's keys) is given.
                from .lpips_vgg import LPIPS
                state = _opt(opt, "lpips_state", None)
                if state is None and not _opt(opt, "allow_random_lpips", False):
                    raise ValueError("lambda_G1_lpips / lambda_G2_lpips > 0 need the lpips package's VGG16 checkpoint as opt.lpips_state "
                                     "(lpips.LPIPS(net='vgg').state_dict()); without it the perceptual terms would optimise a randomly "
                                     "initialised network.  Benchmarks may pass allow_random_lpips=True.")
                self.lpips = LPIPS(net="vgg").to(self.device)
                if state is not None:
                    res = self.lpips.load_state_dict(state, strict=False)
                    # the ScalingLayer constants are buffers with fixed values and `lins.N` aliases `linN`: only a missing VGG16
                    # conv or `linN` weight leaves part of the criterion randomly initialised
                    missing = [k for k in res.missing_keys if not (k.startswith("scaling_layer.") or k.startswith("lins."))]
                    if missing:
                        res = argparse.Namespace(missing_keys=missing)
                    if missing:
                        raise RuntimeError("opt.lpips_state does not cover the LPIPS-VGG16 parameters; missing keys: %s"
                                           % ", ".join(res.missing_keys[:8]))
                self.lpips.refresh_packs_once()
            self.nce_layers = [int(i) for i in str(_opt(opt, "nce_layers", "0,4,8,12,16")).split(",")] if _opt(opt, "lambda_NCE", 0.0) > 0 else []
            if self.nce_layers:
                if not isinstance(self.netG, networks.ResnetGenerator):
                    raise NotImplementedError("PatchNCE needs a generator with feature taps (forward(layers=..., encode_only=True)): the resnet family "
                                              "(the reference's U-Net generators do not support it either, networks.py:1538)")
                if opt.netF not in ("sample", "mlp_sample"):
                    raise NotImplementedError("train-step PatchNCE wiring is built for netF='sample' / 'mlp_sample' (PatchSampleF)")
                self.netF = networks.define_F(g_in, opt.netF, opt.normG, not opt.no_dropout, opt.init_type, opt.init_gain,
                                              opt.no_antialias, gpu, opt)
                if self.netF.use_mlp:
                    self.model_names = self.model_names + ["F"]
                bad = [i for i in self.nce_layers if i not in self.netG.tappable_layers()]
                if bad:
                    raise NotImplementedError("nce_layers %s are not exposed by the fused generator (available: %s)" % (bad, sorted(self.netG.tappable_layers())))
            self.loss_buf = torch.zeros(16, dtype=torch.float32, device=self.device)
        self._spe_cache = {}
        self._losses = {}
        self._graph = None          # captured train step (torch.cuda.CUDAGraph) + the attributes it produced
        self._graph_attrs = None
        self._graph_key = None
        self._input_gen = 0         # bumped whenever an input buffer is (re)allocated -> invalidates the graph

    # ------------------------------------------------------------------ data staging
    def _spe(self, n, h, w):
        key = (n, h, w)
        if key not in self._spe_cache:
            self._spe_cache[key] = spe_grid(h, w, 4, n).to(self.device)
        return self._spe_cache[key]

    def set_input(self, input, phase="train", timing=False, verbose=False):
        """Host tensors (the dataset dict, sinskitG_model.py:702-793) -> persistent device buffers: one pinned H2D copy
        per tensor, then the masking of S / I / T (:724,734,789-790) as in-place device kernels.  `timing` / `verbose` are the
        reference's logging switches (accepted, nothing to print)."""
        dev = self.device
        opt = self.opt
        self.data_phase = phase
        n, _, h, w = input["S"].shape
        # without use_bg_mask the reference skips every mask multiply (:721-726,1317-1319,1339-1341): an all-ones mask is the same
        M = input["M"].float() if opt.use_bg_mask else torch.ones(n, 1, h, w, device=input["S"].device)
        host = {"M": M, "real_S": input["S"].float()}
        if "I" in input:
            host["real_I"] = input["I"].float()
        pre = "" if phase == "train" else "val_"
        has_T = self.isTrain and (pre + "T_images") in input and len(input[pre + "T_images"]) > 0
        if has_T:
            T = input[pre + "T_images"].float()
            NT = T.shape[1]
            if tuple(T.shape[-2:]) != (32, 32) or T.shape[2] != 2:
                raise NotImplementedError("touch patches must be [N, NT, 2, 32, 32] (patch_crop_size 32, T_resolution_multiplier 1); got %s"
                                          % (tuple(T.shape),))
            host["I_masks"] = input[pre + "I_masks"].float().reshape(NT, 1, 32, 32)
            host["real_T"] = T.reshape(NT, 2, 32, 32)
            ox, oy, cs = find_coords_for_patch(input[pre + "T_coords"].cpu().numpy() if torch.is_tensor(input[pre + "T_coords"]) else input[pre + "T_coords"])
            if np.any(np.asarray(cs) != 32 * opt.T_resolution_multiplier):
                # the reference then cuts a smaller / larger window and resizes it (model_utils.py:337-341): not built
                raise NotImplementedError("patch cutout size %s != patch size 32 (resize_ratio != 1): the resampling gather of "
                                          "get_patch_in_input is not built on the B200 path" % sorted(set(np.asarray(cs).tolist())))
            host["ox"] = torch.from_numpy(ox.astype(np.int32))
            host["oy"] = torch.from_numpy(oy.astype(np.int32))
            self.NT = NT
        sc = self._style_code_from(input) if _opt(opt, "use_style_code", False) else None
        if sc is not None:
            host["style_code"] = sc
        self.h2d_bytes = 0
        for k, v in host.items():
            v = v.contiguous()
            if v.is_cuda:
                # items of the device dataset (data_pipeline.SingleSkitDataset) are already in HBM: a device-to-device copy into the
                # persistent buffer (the source is the dataset's cached tensor and must not be masked in place)
                self._stage(k, v.to(dev))
                continue
            if not v.is_pinned():
                v = v.pin_memory()
            self.h2d_bytes += v.numel() * v.element_size()
            self._stage(k, v)
        if sc is None:
            self.style_code = None
        self.M_T = self.M      # mask of the touch output: F.interpolate(M, size * T_resolution_multiplier) with multiplier 1 (:725)
        # background / contact masking on the device (the reference does it after its own .to(device): :724,734,789-790)
        if opt.use_bg_mask:
            ops.mask_mul_(self.real_S, self.M)
            if "I" in input:
                ops.mask_mul_(self.real_I, self.M)
        if has_T:
            ops.mask_mul_(self.real_T, self.I_masks)
        self.S_pe = self._spe(n, h, w) if opt.use_positional_encoding else None
        if self.isTrain and has_T:
            NT, NF = self.NT, opt.add_fake_T_sample_size
            if getattr(self, "fake_in", None) is None or self.fake_in.shape[0] != NT or self.more_in.shape[0] != NF:
                self.fake_in = torch.zeros(NT, 7, 32, 32, device=dev)
                self.real_in = torch.zeros(NT, 7, 32, 32, device=dev)
                self.more_in = torch.ones(NF, 7, 32, 32, device=dev)
                self._input_gen += 1
            self.fake_in[:, 6:7] = self.I_masks
            self.real_in[:, 6:7] = self.I_masks
            self.real_in[:, 0:2] = self.real_T
            self._offset_table = self._make_offset_table(input, M) if opt.use_more_fakeT else None
        self.name = input.get("name")
        self.image_paths = input.get("S_paths", [])
        self.augmentation_params = input.get("augmentation_params")
        self.full_T_coords = input.get("full_T_coords")

    def _make_offset_table(self, input, M):
        """Candidate offsets of the NF random fake patches (model_utils.py:212-218) without stalling the launching stream: items of the
        device dataset carry the table (built when the item was made); a host mask is uploaded and dilated on a side stream of its
        own, so the small read-back waits for that stream only, not for the previous train step still running on the main one."""
        if "M_box_bits" in input:
            bits, rc = input["M_box_bits"], input["M_box_rowcount"]
            bits = bits.cpu().numpy() if torch.is_tensor(bits) else np.asarray(bits)
            rc = rc.cpu().numpy() if torch.is_tensor(rc) else np.asarray(rc)
            return offset_table_from_arrays(bits, rc, M.shape[-2], M.shape[-1])
        if M.is_cuda:
            return random_patch_offset_table(M)
        if bool((M[0, 0] > 0).all()):
            return random_patch_offset_table(M)          # every position is a candidate: nothing to compute
        if getattr(self, "_aux", None) is None:
            self._aux = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(self._aux):
            md = M[:1].contiguous().pin_memory().to(self.device, non_blocking=True)
            return random_patch_offset_table(md)

    def _style_code_from(self, input):
        """The generator's style code as a host fp32 [N, style_code_dim] tensor.  The reference computes it in forward() with CLIP
        ViT-B/32 in fp16 (skitG_model.py:483-489,705-724,1294-1298) and torch.cat promotes it to fp32 inside the U-Net; CLIP is a
        third-party network that is absent here, so the code is taken precomputed (`input['style_code']`) or from a
        user-supplied encoder (`self.style_encoder(style_I or real_I) -> [N, dim]`)."""
        if input.get("style_code") is not None:
            sc = input["style_code"]
        elif getattr(self, "style_encoder", None) is not None:
            img = input.get("style_I", input.get("I"))
            if img is not None and "style_M" in input and self.opt.use_bg_mask:
                img = img * input["style_M"]
            with torch.no_grad():
                sc = self.style_encoder(img)
        else:
            raise RuntimeError("use_style_code is on but the batch carries no 'style_code' and the model has no style_encoder "
                               "(the reference's CLIP ViT-B/32 image encoder is a third-party network outside the B200 path)")
        sc = torch.as_tensor(sc).detach().to("cpu", torch.float32)
        if sc.dim() != 2 or sc.shape[1] != _opt(self.opt, "style_code_dim", 512):
            raise ValueError("style_code must be [N, %d], got %s" % (_opt(self.opt, "style_code_dim", 512), tuple(sc.shape)))
        return sc

    def _stage(self, name, host_t):
        """Host tensor -> a persistent device buffer of the same name (same address every step, so the
        captured CUDA graph of the train step stays valid); reallocates only when the shape changes."""
        cur = getattr(self, name, None)
        if torch.is_tensor(cur) and cur.is_cuda and cur.shape == host_t.shape and cur.dtype == host_t.dtype:
            cur.copy_(host_t, non_blocking=True)
        else:
            # a tensor that already lives on the device is cloned: the staged buffers are masked in place below
            setattr(self, name, host_t.to(self.device).clone() if host_t.is_cuda else host_t.to(self.device, non_blocking=True))
            self._input_gen += 1

    _RING = 4

    def _stage_rand(self, rand):
        """The step's host-side random draws (DiffAugment torch.rand x4, NF random-patch offsets) and the
        Adam scalars go into ONE pinned blob and one H2D copy into a persistent device blob, before the
        (possibly replayed) step body.  A small ring of pinned blobs, each guarded by an event, keeps the
        host from overwriting a blob whose copy the GPU has not executed yet (the host runs ahead of a
        replayed graph)."""
        opt = self.opt
        n = self.real_S.shape[0]
        NF = opt.add_fake_T_sample_size if self.isTrain else 0
        words = 4 * n + 2 * max(NF, 1) + 8
        if getattr(self, "_blob_dev", None) is None or self._blob_dev.numel() != words:
            self._blob_host = [torch.zeros(words, dtype=torch.float32).pin_memory() for _ in range(self._RING)]
            self._blob_evt = [None] * self._RING
            self._blob_i = 0
            self._blob_dev = torch.zeros(words, dtype=torch.float32, device=self.device)
            self._u_dev = self._blob_dev[:4 * n].view(4, n)
            self._fo_dev = self._blob_dev[4 * n:4 * n + 2 * max(NF, 1)].view(torch.int32).view(2, max(NF, 1))
            self._hy_dev = self._blob_dev[4 * n + 2 * max(NF, 1):].view(4, 2)
            self._input_gen += 1
        i = self._blob_i
        self._blob_i = (i + 1) % self._RING
        if self._blob_evt[i] is not None:
            self._blob_evt[i].synchronize()
        h = self._blob_host[i]
        hu = h[:4 * n].view(4, n)
        hfo = h[4 * n:4 * n + 2 * max(NF, 1)].view(torch.int32).view(2, max(NF, 1))
        hhy = h[4 * n + 2 * max(NF, 1):].view(4, 2)
        if opt.use_diffaug:
            for j, k in enumerate(("real_b", "real_s", "fake_b", "fake_s")):
                # the reference's torch.rand draws, in its order: real (b, s) then fake (b, s)
                hu[j].copy_(torch.rand(n, 1, 1, 1).reshape(n) if rand is None else
                            torch.as_tensor(np.asarray(rand[k], dtype=np.float32)).reshape(n))
        if self.isTrain and opt.use_more_fakeT and NF:
            if rand is not None and "fake_ox" in rand:
                fox, foy = np.asarray(rand["fake_ox"], dtype=np.int32), np.asarray(rand["fake_oy"], dtype=np.int32)
            else:
                fox, foy = self._offset_table.sample(NF)
            hfo[0].copy_(torch.from_numpy(np.ascontiguousarray(fox)))
            hfo[1].copy_(torch.from_numpy(np.ascontiguousarray(foy)))
        if self.isTrain:
            t = max(1, self.step_count)
            bc1 = 1.0 - opt.beta1 ** t
            bc2 = 1.0 - opt.beta2 ** t
            for j, lr in enumerate((opt.lr, opt.lr_G2, opt.lr, opt.lr)):   # rows: D, D2, G, F
                hhy[j, 0] = lr * self.lr_factor / bc1
                hhy[j, 1] = 1.0 / math.sqrt(bc2)
        self._blob_dev.copy_(h, non_blocking=True)
        if self.isTrain and getattr(self, "nce_layers", None):
            # PatchSampleF ids (networks.py:703-705: np.random.permutation(H*W)[:num_patches], shared over the batch)
            S_h, S_w = self.real_S.shape[2:]
            P = opt.num_patches
            if getattr(self, "_ids_dev", None) is None or self._ids_dev.shape != (len(self.nce_layers), P):
                self._ids_host = [torch.zeros(len(self.nce_layers), P, dtype=torch.int32).pin_memory() for _ in range(self._RING)]
                self._ids_dev = torch.zeros(len(self.nce_layers), P, dtype=torch.int32, device=self.device)
                self._input_gen += 1
            hi = self._ids_host[i]
            self._ids_count = []
            for li, l in enumerate(self.nce_layers):
                fh, fw = self.netG.feature_hw(l, S_h, S_w)
                if rand is not None and "nce_ids" in rand:
                    ids = np.asarray(rand["nce_ids"][li], dtype=np.int64)
                else:
                    ids = networks.first_of_permutation(fh * fw, P)
                self._ids_count.append(len(ids))
                hi[li, :len(ids)].copy_(torch.from_numpy(ids.astype(np.int32)))
            self._ids_dev.copy_(hi, non_blocking=True)
        evt = torch.cuda.Event()
        evt.record()
        self._blob_evt[i] = evt

    # ------------------------------------------------------------------ forward
    def forward(self, save=None, rand=None, staged=False, real_aug=True):
        """sinskitG_model.py:1293-1344: G, channel split, *M, normal, DiffAugment('bs') real + fake, *M."""
        opt = self.opt
        save = self.isTrain if save is None else save
        srcs = [self.real_S] + ([self.S_pe] if self.S_pe is not None else [])
        taps = set(self.nce_layers) if (save and getattr(self, "nce_layers", None)) else None
        (self.fake_I, self.fake_T, self.fake_N), self._g_ctx, self._g_feats = self.netG.fwd(
            srcs, mask=self.M if opt.use_bg_mask else None, scale_nz=opt.scale_nz, save=save, taps=taps,
            style_code=self.style_code if getattr(opt, "use_style_code", False) else None)
        if hasattr(self, "real_I"):
            if opt.use_diffaug:
                if not staged:
                    self._stage_rand(rand)
                u = self._u_dev
                if real_aug:    # the train step computes it on the real-data branch instead
                    self.aug_real_I = ops.diffaug_bs_mask(self.real_I, self.M, u[0], u[1])
                self.aug_fake_I = ops.diffaug_bs_mask(self.fake_I, self.M, u[2], u[3])
            else:
                self.aug_real_I, self.aug_fake_I = self.real_I, self.fake_I
        return self.fake_I, self.fake_T, self.fake_N

    def test(self, timing=False):
        """sinskitG_model.py:795-807: forward without saving anything for backward (`timing`: the reference's print switch).  With opt.cuda_graph the forward is
        captured once per input shape and replayed (inputs live in the persistent buffers set_input fills)."""
        self.netG.ensure_flat()
        key = (self._input_gen, tuple(self.real_S.shape), self.netG.flat_param.data_ptr())
        if getattr(self, "_tgraph", None) is not None and self._tgraph_key == key:
            self._tgraph.replay()
            L_.launches += self._tgraph_launches
            self.__dict__.update(self._tgraph_attrs)
            return self.fake_I, self.fake_T, self.fake_N
        self._tgraph = None
        self.netG.refresh_packs()
        self._test_calls = getattr(self, "_test_calls", 0) + 1
        if _opt(self.opt, "cuda_graph", True) and self._test_calls > 1:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            before = dict(self.__dict__)
            l0 = L_.launches
            with torch.cuda.graph(g):
                self.forward(save=False, staged=True)
            self._tgraph_launches = L_.launches - l0
            self._tgraph_attrs = {k: v for k, v in self.__dict__.items() if k not in before or before[k] is not v}
            self._tgraph, self._tgraph_key = g, key
            for k in ("_tgraph", "_tgraph_key", "_tgraph_attrs", "_tgraph_launches"):
                self._tgraph_attrs.pop(k, None)
            g.replay()
            return self.fake_I, self.fake_T, self.fake_N
        return self.forward(save=False)

    # ------------------------------------------------------------------ helpers of the step
    def _allreduce(self, net):
        if self.dist is not None:
            self.dist.allreduce_grads(net.flat_grad)

    def _adam(self, net, slot):
        """slot: row of the device hyper buffer (0 = D, 1 = D2, 2 = G) written by _stage_rand."""
        scale = 1.0 / self.dist.world_size if self.dist is not None else 1.0
        ops.adam_step_dev(net.flat_param, net.flat_grad, net.exp_avg, net.exp_avg_sq, self._hy_dev[slot],
                          self.opt.beta1, self.opt.beta2, 1e-8, scale)
        net.refresh_packs()

    def _gan(self, preds, sign, loss_slot, gscale=None):
        """Sum over scales of the per-sample GAN loss (softplus or hinge; sign -1 = real target, +1 = fake); returns the per-scale
        dpred list if gscale."""
        mode = _opt(self.opt, "gan_mode", "nonsaturating")
        dps = []
        for p in preds:
            dp = torch.empty_like(p) if gscale is not None else None
            ops.gan_loss(p, mode, sign < 0, 0.0, loss_slot, dp, gscale or 0.0)
            dps.append(dp)
        return dps

    # ------------------------------------------------------------------ the train step
    def optimize_parameters(self, epoch=None, rand=None):
        """sinskitG_model.py:601-700.  rand (optional, for parity tests): dict with the DiffAugment draws
        real_b/real_s/fake_b/fake_s and the NF random fake-patch offsets fake_ox/fake_oy; drawn from
        torch / random like the reference when absent."""
        opt = self.opt
        G, D, D2 = self.netG, self.netD, self.netD2
        self.step_count += 1
        for net in (G, D, D2):
            net.ensure_flat()
        if self.step_count == 1:
            for net in (G, D, D2):
                net.refresh_packs()
        self._stage_rand(rand)
        key = (self._input_gen, tuple(self.real_S.shape), self.NT, opt.add_fake_T_sample_size,
               tuple(net.flat_param.data_ptr() for net in (G, D, D2)),
               getattr(getattr(self, "netF", None), "flat_param", None) is not None)
        if self._graph is not None and self._graph_key == key:
            self._graph.replay()
            L_.launches += self._graph_launches   # kernel-launching ABI calls the replayed graph stands for
            self.__dict__.update(self._graph_attrs)
            return self._loss_raw[0]
        self._graph = None
        if _opt(opt, "cuda_graph", True) and self.step_count > _opt(opt, "cuda_graph_warmup", 2):
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            before = dict(self.__dict__)
            l0 = L_.launches
            with torch.cuda.graph(g):
                self._step_body()
            self._graph_launches = L_.launches - l0
            self._graph, self._graph_key = g, key
            self._graph_attrs = {k: v for k, v in self.__dict__.items() if k not in before or before[k] is not v}
            for k in ("_graph", "_graph_key", "_graph_attrs", "_graph_launches"):
                self._graph_attrs.pop(k, None)
            g.replay()    # capture does not execute: run the step once
            return self._loss_raw[0]
        return self._step_body()

    # -- parallel branches of the step (side streams; parallel branches of the captured graph)
    def _fork(self, i):
        bs = getattr(self, "_bstreams", None)
        if bs is None:
            bs = self._bstreams = [torch.cuda.Stream(device=self.device) for _ in range(7)]
        bs[i].wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(bs[i])

    def _comm_stream(self):
        if getattr(self, "_cstream", None) is None:
            self._cstream = torch.cuda.Stream(device=self.device)
        return self._cstream

    def _join(self, *idx):
        cur = torch.cuda.current_stream()
        for i in idx:
            cur.wait_stream(self._bstreams[i])

    def _d_pass(self, net, srcs, sign, slot, gscale, deferred):
        """One discriminator pass of a D step: forward, softplus GAN loss, full backward (weight gradients accumulate
        atomically into the net's flat bucket, so passes on parallel streams may overlap)."""
        preds, ctx = net.fwd(srcs, deferred=deferred)
        net.bwd(ctx, self._gan(preds, sign, slot, gscale))
        return preds

    def _step_body(self):
        if getattr(self, "_arena", None) is None:
            self._arena = ops.ZeroArena(self.device)
        ops.ARENA = self._arena
        self._arena.begin()
        try:
            return self._step_body_inner()
        finally:
            ops.ARENA = None

    def _step_body_inner(self):
        """All device work of one train step; reads only persistent device buffers (graph-capturable).
        Dependency structure used for overlap: the real-data passes of D and D2 do not depend on the generator, the
        D2 passes do not depend on D, and nothing but the optimiser consumes weight gradients.  BatchNorm running
        statistics are applied after the joins in the reference's order (fake, [full-res], more, real)."""
        opt = self.opt
        G, D, D2 = self.netG, self.netD, self.netD2
        NT, NF = self.NT, opt.add_fake_T_sample_size
        n = self.real_S.shape[0]
        ox, oy = self.ox, self.oy
        L = ops.zeros((8 + 3 * NT + NF,), torch.float32, self.device)      # a slice of the step's zero arena
        sl = dict(D_fake=L[0:1], D_real=L[1:2], G_GAN=L[2:3], G_L1=L[3:4], G2_L1=L[4:5], G_lpips=L[5:6], G2_lpips=L[6:7],
                  D2_fake=L[8:8 + NT], D2_real=L[8 + NT:8 + 2 * NT], G2_GAN=L[8 + 2 * NT:8 + 3 * NT], D2_more=L[8 + 3 * NT:])
        D.zero_grad()
        D2.zero_grad()
        G.zero_grad()        # before any branch: the PatchNCE query branch accumulates encoder weight gradients beside the D steps
        run_D, run_D2 = {}, {}
        more = opt.use_more_fakeT and NF

        # ---- real-data passes: independent of G, start them first
        with self._fork(0):     # D1 real (compute_D1_loss :1346-1407)
            run_D["real"] = []
            self._d_pass(D, [self.real_S, self.real_I], -1.0, sl["D_real"], 0.5 * opt.lambda_G1_GAN / n, run_D["real"])
            if self.lpips is not None:      # VGG features of the real image / real touch patches: independent of G as well
                if opt.lambda_G1_lpips > 0:
                    self._lp_real_I = self.lpips.fwd(self.real_I)
                if opt.lambda_G2_lpips > 0:   # gx patches then gy patches, one channel each (:1639-1645)
                    self._lp_real_T = self.lpips.fwd(self.real_T.transpose(0, 1).reshape(-1, 1, 32, 32))
        with self._fork(1):     # D2 real (compute_D2_loss :1409-1617); its conditioning image is the DiffAugmented real
            if opt.use_diffaug:
                self.aug_real_I = ops.diffaug_bs_mask(self.real_I, self.M, self._u_dev[0], self._u_dev[1])
            else:
                self.aug_real_I = self.real_I
            ops.patch_gather([self.real_S, self.aug_real_I], ox, oy, 32, ctot=7, coffs=[2, 3], dst=self.real_in)
            run_D2["real"] = []
            self._d_pass(D2, [self.real_in], -1.0, sl["D2_real"], 0.5 * opt.lambda_G2_GAN / NT, run_D2["real"])

        # ---- generator forward + compute_additional_output (:1268-1291): patch gathers straight into the D2 inputs
        self.forward(save=True, staged=True, real_aug=False)
        fake_I, fake_T = self.fake_I, self.fake_T
        fake_T_p = ops.patch_gather([fake_T], ox, oy, 32)
        ops.patch_gather([fake_T, self.real_S, self.aug_fake_I], ox, oy, 32, ctot=7, dst=self.fake_in)
        if more:
            ops.patch_gather([fake_T, self.real_S, fake_I], self._fo_dev[0], self._fo_dev[1], 32, ctot=7, dst=self.more_in)

        # ---- perceptual terms on the generated image / patches: they need only G's output and the real features from
        #      branch 0, so the whole VGG forward + input gradient runs on its own branch beside both discriminator
        #      steps; its gradients are added to dI / dTp just before G's backward
        if self.lpips is not None:
            lp = self._lp_out = {}
            if opt.lambda_G1_lpips > 0:      # compute_G1_loss :1707-1715: LPIPS(fake_I, real_I).mean() * lambda
                with self._fork(4):
                    torch.cuda.current_stream().wait_stream(self._bstreams[0])
                    lp["I"] = self.lpips.loss_and_grad(fake_I, None, gscale=opt.lambda_G1_lpips / n, real_feats=self._lp_real_I)
            if opt.lambda_G2_lpips > 0:      # _compute_touch_lpips_loss :1619-1658: sum over the NT patches, mean over images
                with self._fork(5):          # 128 images of 32x32: launch-latency bound, so beside the full-image pass
                    torch.cuda.current_stream().wait_stream(self._bstreams[0])
                    lp["T"] = self.lpips.loss_and_grad(fake_T_p.transpose(0, 1).reshape(-1, 1, 32, 32), None,
                                                       gscale=opt.lambda_G2_lpips / n, real_feats=self._lp_real_T)

        # ---- PatchNCE query branch (encoder forward + backward on the channel mean of fake_I): it needs only G's output and
        #      G's (not yet updated) weights, so it runs beside both discriminator steps; its gradient w.r.t. the query sketch
        #      (one channel) is folded into dI just before G's backward
        if self.nce_layers:
            with self._fork(6):
                self._nce_dSq = self._nce_step(n)

        # ---- fake passes: D2 (patches) and D2 (random patches) beside D1 (full image)
        with self._fork(2):
            run_D2["fake"] = []
            self._d_pass(D2, [self.fake_in], +1.0, sl["D2_fake"], 0.5 * opt.lambda_G2_GAN / NT, run_D2["fake"])
            if _opt(opt, "run_full_res_D2", False):  # visualisation only in the reference (:1495-1500); updates BN running stats
                self.pred_fake_T_full = D2.fwd([fake_T, self.real_S, self.aug_fake_I, self.M], save=False, deferred=run_D2["fake"])[0][-1]
        if more:
            with self._fork(3):
                run_D2["more"] = []
                self._d_pass(D2, [self.more_in], +1.0, sl["D2_more"], 0.5 * opt.lambda_G2_GAN / NF, run_D2["more"])
        run_D["fake"] = []
        self.pred_fake_I = self._d_pass(D, [self.real_S, fake_I], +1.0, sl["D_fake"], 0.5 * opt.lambda_G1_GAN / n, run_D["fake"])[-1]

        # ---- D1 update (:648-653)
        self._join(0)
        for k in ("fake", "real"):
            networks.apply_running_updates(run_D[k])
        self._allreduce(D)
        self._adam(D, 0)

        # ---- D2 update (:654-667) and the G step's value-only D2 pass stay on branch 1, off the launching stream: the
        #      generator's backward does not depend on them (G2 GAN is computed on a detached clone, :1751,1781)
        b1 = self._bstreams[1]
        for i in (2, 3) if more else (2,):
            b1.wait_stream(self._bstreams[i])
        with torch.cuda.stream(b1):
            for k in ("fake", "more", "real"):
                if k in run_D2:
                    networks.apply_running_updates(run_D2[k])
            self._allreduce(D2)
            self._adam(D2, 1)
            pg2, _ = D2.fwd([self.fake_in], save=False)
            self._gan(pg2, -1.0, sl["G2_GAN"])

        # ---- G step (:680-694): GAN through the updated (frozen) D, L1, patch L1
        pg, cg = D.fwd([self.real_S, fake_I])
        dpg = self._gan(pg, -1.0, sl["G_GAN"], opt.lambda_G1_GAN / n)
        dI = D.bwd(cg, dpg, need_wgrad=False, input_slice=(self.sketch_nc, 3))
        del cg
        ops.l1_loss(fake_I, self.real_I, opt.lambda_G1_L1 / fake_I.numel(), sl["G_L1"], dI, opt.lambda_G1_L1 / fake_I.numel(), accumulate=True)
        per_patch = fake_T_p.numel() // NT
        dTp = torch.empty_like(fake_T_p)
        ops.l1_loss(fake_T_p, self.real_T, opt.lambda_G2_L1 / per_patch / n, sl["G2_L1"], dTp, opt.lambda_G2_L1 / per_patch / n)
        if self.lpips is not None:
            self._join(*([4] if "I" in self._lp_out else []), *([5] if "T" in self._lp_out else []))
            if "T" in self._lp_out:
                ln2, dxp = self._lp_out["T"]
                dTp.add_(dxp.view(2, -1, 32, 32).transpose(0, 1))
                sl["G2_lpips"].add_(ln2.sum() * (opt.lambda_G2_lpips / n))
            if "I" in self._lp_out:
                ln, dI_lp = self._lp_out["I"]
                dI.add_(dI_lp)
                sl["G_lpips"].add_(ln.sum() * (opt.lambda_G1_lpips / n))
            self._lp_out = None
        dT = ops.zeros_big(fake_T.shape, torch.float32, self.device)
        ops.patch_scatter_add(dTp, 0, 2, ox, oy, dT)
        if self.nce_layers:
            self._join(6)
            if self._nce_dSq is not None:
                ops.channel_mean_bwd(self._nce_dSq, dI)
            self._nce_dSq = None
            self._g_feats = None      # released on the launching stream, after the branch that read them has joined
        if self.dist is not None and isinstance(G, networks.ResnetGenerator) and G.n_blocks >= 4:
            # data parallel: the gradient bucket's tail (later ResnetBlocks, up-convs, head: about half of the 45.6 MB) is final
            # half-way through the backward pass — its all-reduce starts there on a communication stream, beside the rest of the
            # serial input-gradient chain; only the head of the bucket is reduced after the pass
            split = {}

            def tail_done(off):
                comm = self._comm_stream()
                comm.wait_stream(torch.cuda.current_stream())
                side = networks._wgrad_side_stream()
                if side is not None:
                    comm.wait_stream(side)
                with torch.cuda.stream(comm):
                    self.dist.allreduce_grads(G.flat_grad[off:])
                split["off"] = off

            G.bwd(self._g_ctx, dI, dT, on_tail_done=tail_done, tail_block=G.n_blocks // 2)
            self._g_ctx = None
            self._join(1)
            self.dist.allreduce_grads(G.flat_grad[:split["off"]])
            torch.cuda.current_stream().wait_stream(self._comm_stream())
        else:
            G.bwd(self._g_ctx, dI, dT)
            self._g_ctx = None
            self._join(1)
            self._allreduce(G)
        self._adam(G, 2)
        if self.nce_layers and self.netF.use_mlp:
            self._allreduce(self.netF)
            self._adam(self.netF, 3)
        self._loss_raw = (L, NT, NF)
        return L

    def _nce_step(self, n):
        """CUT-style PatchNCE term (new wiring: PatchNCELoss / PatchSampleF are dead code in the reference).
        keys  = generator features of the input (sketch + positional encoding) at nce_layers — tapped from the main forward;
        query = the same encoder run on the 1-channel mean of the generated RGB image (+ the same positional encoding);
        loss  = lambda_NCE * mean_layers mean_patches PatchNCE(q, k);  its gradient flows into the encoder weights (query
        pass) and, through the channel mean, into fake_I.  Returns dLoss/dSq ([n, 1, H, W]; the caller spreads it over the
        three channels of dI with channel_mean_bwd) or None."""
        opt, G = self.opt, self.netG
        S_h, S_w = self.real_S.shape[2:]
        Sq = ops.channel_mean(self.fake_I)
        fq, cq = G.encode([Sq] + ([self.S_pe] if self.S_pe is not None else []), self.nce_layers)
        nl = len(self.nce_layers)
        dfeats, chunks = {}, []
        F_ = self.netF
        if F_.use_mlp:
            if not F_.mlp_init:   # created on first use like the reference (networks.py:678-686); eager warm-up step
                F_.create_mlp(channels=[int(self._g_feats[l].shape[3]) for l in self.nce_layers], device=self.device)
                F_.flatten_parameters()
                if self.dist is not None:
                    self.dist.broadcast_params([F_])
            F_.ensure_flat()
            if not getattr(F_, "_packed_once", False):
                F_.refresh_packs()
                F_._packed_once = True
            F_.zero_grad()
        for li, l in enumerate(self.nce_layers):
            ids = self._ids_dev[li][:self._ids_count[li]]
            k_pool, _ = F_.sample_fwd(self._g_feats[l], ids, li)
            q_pool, qctx = F_.sample_fwd(fq[l], ids, li, save=True)
            rows = q_pool.shape[0]
            b = 1 if _opt(opt, "nce_includes_all_negatives_from_minibatch", False) else n
            loss_l, dq = ops.patchnce(q_pool, k_pool, b, opt.nce_T, want_grad=True, gscale=opt.lambda_NCE / (nl * rows))
            dfeats[l] = F_.sample_bwd(qctx, dq)
            chunks.append(loss_l)
        dpad0 = G.encode_bwd(cq, dfeats, input_channels=1)      # only the query sketch channel carries a gradient (to fake_I)
        dSq = None
        if dpad0 is not None:
            dSq = ops.operand_grad_to_nchw(dpad0, S_h, S_w, 3, ops.PAD_REFLECT, 0, 1)
        if dfeats.get(0) is not None and (dpad0 is None or dpad0.shape[3] == 1):   # the layer-0 tap is the padded input itself
            dSq = ops.operand_grad_to_nchw(dfeats[0], S_h, S_w, 3, ops.PAD_REFLECT, 0, 1, dst=dSq, accumulate=dSq is not None)
        self._nce_losses = chunks
        return dSq

    def current_losses(self):
        """Device -> host read of the step's loss scalars, bare names (the reference does ~12 .item() syncs per step,
        sinskitG_model.py:1389-1838; here it is one D2H copy, only when somebody asks)."""
        L, NT, NF = self._loss_raw
        v = L.detach().cpu().numpy()
        o = self.opt
        extra = {}
        if getattr(self, "nce_layers", None) and getattr(self, "_nce_losses", None):
            per_layer = torch.stack([c.mean() for c in self._nce_losses]).cpu().numpy()
            extra["NCE"] = float(per_layer.mean()) * o.lambda_NCE
        if getattr(self, "lpips", None) is not None:
            if o.lambda_G1_lpips > 0:
                extra["G_lpips"] = float(v[5])
            if o.lambda_G2_lpips > 0:
                extra["G2_lpips"] = float(v[6])
        return dict(extra,
            D_fake_I=float(v[0]) * o.lambda_G1_GAN, D_real_I=float(v[1]) * o.lambda_G1_GAN,
            G_GAN=float(v[2]) * o.lambda_G1_GAN, G_L1=float(v[3]), G2_L1=float(v[4]),
            D_fake_T_concat=float(v[8:8 + NT].mean()) * o.lambda_G2_GAN,
            D_real_T_concat=float(v[8 + NT:8 + 2 * NT].mean()) * o.lambda_G2_GAN,
            G2_GAN=float(v[8 + 2 * NT:8 + 3 * NT].sum()) * o.lambda_G2_GAN,
            D_more_fake_T=float(v[8 + 3 * NT:].mean()) * o.lambda_G2_GAN if NF else 0.0,
        )

    def get_current_losses(self):
        """BaseModel.get_current_losses (base_model.py:171-183): OrderedDict keyed 'l_' + name, in the reference model's
        loss_names order (sinskitG_model.py:430-456); the gradient penalties are identically 0 outside gan_mode 'wgangp'."""
        raw = self.current_losses()
        order = ["G_GAN", "D_real_I", "D_fake_I", "D_I_grad_penalty", "G_L1", "G_lpips", "G2_GAN", "D_real_T_concat",
                 "D_fake_T_concat", "D_T_grad_penalty", "D_more_fake_T", "G2_L1", "G2_lpips", "G2_GAN_feat", "NCE"]
        raw.setdefault("D_I_grad_penalty", 0.0)
        raw.setdefault("D_T_grad_penalty", 0.0)
        if _opt(self.opt, "lambda_G2_GAN_feat", 0.0) > 0.0:
            # listed by the reference whenever the weight is positive (default 1, :455-456) and always 0: its feature-matching branch
            # compares a module with a string (`self.netD2 == "multiscale"`, :1794) and never runs
            raw.setdefault("G2_GAN_feat", 0.0)
        return collections.OrderedDict(("l_" + k, raw[k]) for k in order if k in raw)

    # ------------------------------------------------------------------ BaseModel contract (models/base_model.py:71-230)
    def setup(self, opt=None):
        """base_model.py:90-102: (re)start the LR schedule, load networks for test / continue_train, print the networks."""
        opt = self.opt if opt is None else opt
        self._sched_count = 0
        self.lr_factor = self._lambda_rule(0) if self.isTrain else 1.0
        if not self.isTrain or _opt(opt, "continue_train", False):
            self.load_networks(_opt(opt, "epoch", "latest"))
        self.print_networks(_opt(opt, "verbose", False))

    def parallelize(self):
        """base_model.py:104-108 wraps every net in nn.DataParallel.  The B200 path is one process per GPU (torchrun) with a
        flat-bucket NCCL all-reduce (dist.py), so there is nothing to wrap: with a DistContext the replicas are made
        identical here instead."""
        if self.dist is not None:
            self.dist.broadcast_params([getattr(self, "net" + n) for n in self.model_names if getattr(self, "net" + n, None) is not None
                                        and getattr(getattr(self, "net" + n), "flat_param", None) is not None])

    def data_dependent_initialize(self, data):
        pass

    def _nets(self):
        return [getattr(self, "net" + n) for n in self.model_names if isinstance(n, str) and getattr(self, "net" + n, None) is not None]

    def train(self):
        for net in self._nets():
            net.train()

    def eval(self):
        for net in self._nets():
            net.eval()

    def compute_visuals(self):
        """sinskitG_model.py:1320-1324: the two height-gradient channels as separate maps."""
        if getattr(self, "fake_T", None) is not None:
            self.fake_gx, self.fake_gy = self.fake_T[:, 0:1], self.fake_T[:, 1:2]

    visual_names = ["real_S", "M", "real_I", "fake_I", "fake_gx", "fake_gy", "fake_N", "pred_fake_I", "pred_fake_T_full",
                    "aug_fake_I", "aug_real_I"]

    def get_current_visuals(self):
        """The self-attribute visuals of sinskitG_model.py:811-831 (NCHW device tensors).  The patch collages and the
        evaluation metrics the reference computes inside this call (compute_additional_visuals / compute_evaluation_metric:
        SIFID, LPIPS, PSNR/SSIM — SURVEY.md section 8f rank 4) are not part of the hot path and are not produced."""
        self.compute_visuals()
        ret = collections.OrderedDict()
        for name in self.visual_names:
            v = getattr(self, name, None)
            if torch.is_tensor(v):
                ret[name] = v.permute(0, 3, 1, 2) if name.startswith("pred_") else v    # predictions are stored NHWC
        return ret

    def get_image_paths(self):
        return self.image_paths

    eval_metrics = ["I_PSNR", "I_SSIM", "T_AE", "T_MSE"]     # the built subset of sinskitG_model.py:461-470 (eval_metrics.py)

    def compute_current_metrics(self, prefix=None):
        """The evaluation metrics of the current batch (what compute_evaluation_metric returns inside the reference's
        get_current_visuals, sinskitG_model.py:839-1040, for the samples at hand): PSNR / SSIM of fake_I against real_I, angle
        error / MSE of the generated touch patches against the real ones — device-side reductions, one D2H read.  Sets
        self.metric_<prefix><name> and registers the names for get_current_metrics().  prefix: 'train_' in training, '' otherwise."""
        from .eval_metrics import compute_evaluation_metric
        if prefix is None:
            prefix = "train_" if (self.isTrain and getattr(self, "data_phase", "train") == "train") else ""
        if getattr(self, "real_I", None) is None or getattr(self, "fake_I", None) is None:
            raise RuntimeError("compute_current_metrics needs a batch with ground truth and a forward pass (set_input + forward / test)")
        rT = fT = None
        if getattr(self, "real_T", None) is not None and getattr(self, "fake_T", None) is not None and hasattr(self, "ox"):
            rT, fT = self.real_T, ops.patch_gather([self.fake_T], self.ox, self.oy, 32)
        names = [m for m in self.eval_metrics if rT is not None or not m.startswith("T_")]
        res = compute_evaluation_metric(self.model_names, self.real_I, self.fake_I, rT, fT, eval_metrics=names, prefix=prefix)
        for k, v in res.items():
            setattr(self, k, float(v))
            if k[len("metric_"):] not in self.metric_names:
                self.metric_names.append(k[len("metric_"):])
        return res

    def get_current_metrics(self):
        """base_model.py:185-200: OrderedDict 'm_' + name of whatever compute_current_metrics() has produced so far (the SIFID /
        LPIPS-alex metrics of the reference's list need third-party networks and are not produced)."""
        return collections.OrderedDict(("m_" + n, float(getattr(self, "metric_" + n))) for n in self.metric_names)

    def generate_visuals_for_evaluation(self, data, mode):
        return {}

    def print_networks(self, verbose=False):
        print("---------- Networks initialized -------------")
        for name, net in zip([n for n in self.model_names if getattr(self, "net" + n, None) is not None], self._nets()):
            if verbose:
                print(net)
            print("[Network %s] Total number of parameters : %.3f M" % (name, sum(p.numel() for p in net.parameters()) / 1e6))
        print("-----------------------------------------------")

    def set_requires_grad(self, nets, requires_grad=False):
        """base_model.py:324-335.  The explicit step decides per pass which gradients it computes (need_wgrad / dgrad-only
        discriminator passes); the flags are still set for callers that inspect them."""
        for net in nets if isinstance(nets, list) else [nets]:
            if net is not None:
                for p in net.parameters():
                    p.requires_grad = requires_grad

    def _lambda_rule(self, count):
        o = self.opt
        if o.lr_policy != "linear":
            raise NotImplementedError("lr_policy %r: the fused Adam step takes the 'linear' schedule (the model default)" % o.lr_policy)
        return 1.0 - max(0, count + o.epoch_count - o.n_epochs) / float(o.n_epochs_decay + 1)

    def update_learning_rate(self, epoch=None):
        """BaseModel.update_learning_rate() (base_model.py:145-158): one LambdaLR step per call, i.e. after k calls the factor is
        lambda_rule(k) (networks.py:161-165; the scheduler's counter starts at 0 in setup()).  `epoch` is not part of the
        reference's signature: when given it SETS the counter (resuming), it is not train.py's 1-based epoch."""
        self._sched_count = getattr(self, "_sched_count", 0) + 1 if epoch is None else int(epoch)
        self.lr_factor = self._lambda_rule(self._sched_count)
        print("learning rate = %.7f" % (self.opt.lr * self.lr_factor))

    # ------------------------------------------------------------------ checkpoints (base_model.py:205-304)
    def save_networks(self, epoch):
        os.makedirs(self.save_dir, exist_ok=True)
        for name in self.model_names:
            net = getattr(self, "net" + name, None)
            if net is not None:
                torch.save(collections.OrderedDict((k, v.detach().cpu().clone()) for k, v in net.state_dict().items()),
                           os.path.join(self.save_dir, "%s_net_%s.pth" % (epoch, name)))

    def load_networks(self, epoch):
        """base_model.py:247-304: '<epoch>_net_<name>.pth' from save_dir (or checkpoints_dir/pretrained_name when training),
        'module.' prefixes stripped; a missing file is skipped with a warning like the reference (:264-267), a state_dict
        that does not fit raises (the reference only prints)."""
        o = self.opt
        d = self.save_dir
        if self.isTrain and _opt(o, "pretrained_name", None) is not None:
            d = os.path.join(o.checkpoints_dir, o.pretrained_name)
        for name in self.model_names:
            net = getattr(self, "net" + name, None)
            path = os.path.join(d, "%s_net_%s.pth" % (epoch, name))
            if net is None:
                continue
            if not os.path.exists(path):
                warnings.warn("cannot find model path %s, skip" % path)
                continue
            sd = torch.load(path, map_location="cpu")
            sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
            net.load_state_dict(sd)
            net.refresh_packs()


class SKITGModel(SinSKITGModel):
    """skitG: the multi-material twin (models/skitG_model.py).  Differences from sinskitG that reach the hot path:
      * a 512-d style code per sample enters the U-Net generators, tiled and concatenated at the innermost decoder level(s)
        (networks.py:1600-1633; skitG_model.py:1294-1302) — the resnet generators accept and ignore it (networks.py:1131).
        The reference computes it with CLIP ViT-B/32 in fp16 inside forward(); here it comes precomputed in the batch
        (`style_code`, any float dtype: torch.cat promotes CLIP's fp16 to fp32 in the reference too) or from `style_encoder`.
        It is staged into a persistent device buffer like every other input, so the captured step graph reads the current
        material's code on every replay;
      * `M_T`, the mask of the touch output (skitG_model.py:687,1313), equals M at T_resolution_multiplier = 1;
      * the reference's own skitG.optimize_parameters raises TypeError as shipped: the step is sinskitG's.
    `model_defaults` are skitG's set_defaults (skitG_model.py:296-319)."""
    model_name = "skitG"

    def __init__(self, opt, dist_ctx=None, style_encoder=None):
        self.style_encoder = style_encoder
        super().__init__(opt, dist_ctx)
