"""visual-tactile-synthesis_b200 — B200-native (sm_100a) skitG / sinskitG hot path.

Import as `vts_b200` (the directory name is not a valid Python identifier; `vts_b200.py` at the
repo root registers this package under that name).  Layout:

  csrc/         CUDA kernels + the C ABI (include/skit_b200.h) -> csrc/libskit_b200.so
  _lib.py       ctypes binding (fails loudly when the library is missing)
  ops.py        tensor-level wrappers, one per C entry point
  networks.py   define_G / define_D / define_F / GANLoss / PatchSampleF  (reference: models/networks.py)
  patchnce.py   PatchNCELoss                                             (reference: models/patchnce.py)
  model_utils.py get_patch_in_input / find_coords_for_patch / compute_normal (reference: models/model_utils.py)
  skit_model.py SinSKITGModel / SKITGModel train step + forward           (reference: models/{sinskitG,skitG}_model.py)
  dist.py       one-process-per-GPU data parallel: flat gradient buckets + NCCL all-reduce
  eval_metrics.py compute_evaluation_metric: PSNR / SSIM / angle error / MSE reductions    (reference: models/model_utils.py:431-561)
  data_pipeline.py SingleSkitDataset on the device: Pillow-exact resize, crop, contact-centre search (reference: data/singleskit_dataset.py, data/dataset_util.py)
"""
from . import _lib, ops, networks, model_utils, patchnce, skit_model, dist, sg2_generator, lpips_vgg, eval_metrics, data_pipeline  # noqa: F401
from .networks import define_D, define_F, define_G, GANLoss, PatchSampleF  # noqa: F401
from .patchnce import PatchNCELoss  # noqa: F401
from .model_utils import get_patch_in_input, compute_normal, find_coords_for_patch  # noqa: F401
from .data_pipeline import SingleSkitDataset, SkitDataset  # noqa: F401
from .skit_model import SinSKITGModel, SKITGModel, default_options, reference_default_options  # noqa: F401

__all__ = ["define_G", "define_D", "define_F", "GANLoss", "PatchSampleF", "ops", "networks"]
