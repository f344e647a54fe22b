"""ctypes binding of libskit_b200.so (include/skit_b200.h).

The product path has no CPU fallback: if the library is missing `load()` raises, and every
wrapper raises RuntimeError with the library's own message when a call fails.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SKIT_B200_LIB") or os.path.join(_HERE, "csrc", "libskit_b200.so")   # override: A/B a previous build

FMT_F32, FMT_BF16X2 = 0, 1
PAD_ZERO, PAD_REFLECT, PAD_REPLICATE = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
NORM_NONE, NORM_INSTANCE, NORM_BATCH = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2


class SkitOperand(C.Structure):
    _fields_ = [("p0", C.c_void_p), ("p1", C.c_void_p), ("fmt", C.c_int),
                ("n", C.c_int), ("hp", C.c_int), ("wp", C.c_int), ("c", C.c_int)]


class SkitPackDesc(C.Structure):
    _fields_ = [("w", C.c_void_p), ("f32", C.c_void_p), ("hi", C.c_void_p), ("lo", C.c_void_p), ("start", C.c_longlong),
                ("co", C.c_int), ("ci", C.c_int), ("k", C.c_int), ("mode", C.c_int), ("kpad", C.c_int), ("reserved", C.c_int)]


class SkitWeights(C.Structure):
    _fields_ = [("f32", C.c_void_p), ("hi", C.c_void_p), ("lo", C.c_void_p),
                ("k", C.c_int), ("ci", C.c_int), ("co", C.c_int), ("kw", C.c_int)]


_P, _I, _F, _D, _LL = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_longlong
_OP, _WT = C.POINTER(SkitOperand), C.POINTER(SkitWeights)

# name -> argtypes; must list every symbol include/skit_b200.h declares (tests/test_abi.py checks).
SIGNATURES = {
    "skit_pack_conv_weights": [_P, _I, _I, _I, _I, _P, _P, _P, _P],
    "skit_pack_conv_weights_padded": [_P, _I, _I, _I, _I, _I, _P, _P, _P],
    "skit_conv2d_wgrad_ex": [_OP, _I, _OP, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _P],
    "skit_pack_conv_weights_batched": [_P, _I, _LL, _P],
    "skit_pack_conv_weights_tiled": [_P, _I, _I, _I, _P],
    "skit_lpips_scale_fwd": [_P, _I, _I, _I, _I, _OP, _P],
    "skit_lpips_scale_bwd": [_P, _I, _I, _I, _I, _F, _P, _I, _I, _I, _P],
    "skit_maxpool2_fwd": [_P, _I, _I, _I, _I, _OP, _I, _P],
    "skit_maxpool2_bwd": [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "skit_lpips_layer": [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P],
    "skit_sg2_weight_prep_bwd": [_P, _I, _I, _I, _I, _P, _P, _P],
    "skit_sg2_bias_act_bwd": [_P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _F, _F, _P, _I, _P, _I, _P, _OP, _I, _P, _P, _P, _P],
    "skit_sg2_weight_prep": [_P, _I, _I, _I, _I, _P, _P],
    "skit_sg2_bias_act": [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _F, _F, _P, _OP, _I, _P, _I, _P],
    "skit_pack_conv_weights_folded": [_P, _I, _I, _I, _I, _I, _P, _P, _P],
    "skit_fold_x_operand": [_OP, _I, _OP, _P],
    "skit_conv2d_wgrad_folded": [_OP, _I, _OP, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P],
    "skit_conv2d_wgrad_dyfolded": [_OP, _OP, _I, _I, _I, _I, _P, _P, _I, _I, _P],
    "skit_dbias_n": [_OP, _I, _I, _I, _I, _P, _P],
    "skit_unpack_conv_wgrad": [_P, _I, _I, _I, _P, _I, _P],
    "skit_conv2d_fwd": [_OP, _WT, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P],
    "skit_conv2d_dgrad_gather": [_P, _I, _I, _I, _I, _WT, _I, _I, _I, _P, _P],
    "skit_conv2d_dgrad_s1": [_OP, _WT, _P, _P],
    "skit_set_backward_terms": [_I],
    "skit_conv2d_dgrad_s2": [_OP, _I, _WT, _I, _I, _I, _I, _I, _P, _P],
    "skit_conv2d_wgrad": [_OP, _I, _OP, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P],
    "skit_stats_finalize": [_P, _I, _I, _D, _F, _P, _P, _P, _F, _P],
    "skit_norm_act_pad": [_P, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P, _P, _OP, _I, _I, _P],
    "skit_norm_act_pad_stats": [_P, _I, _I, _I, _I, _P, _D, _F, _P, _I, _P, _P, _I, _P, _P, _OP, _I, _I, _I, _P],
    "skit_norm_act_pad_ex": [_P, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P, _P, _OP, _I, _I, _I, _P],
    "skit_act_norm_bwd_reduce_ex": [_P, _I, _I, _P, _P, _I, _I, _I, _P, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P],
    "skit_act_norm_bwd_reduce_ex2": [_P, _I, _I, _P, _P, _I, _I, _I, _P, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P],
    "skit_conv_transpose2d_fwd": [_P, _I, _I, _I, _I, _WT, _I, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P],
    "skit_dbias": [_OP, _I, _I, _I, _P, _P],
    "skit_g_head_bwd_split": [_P, _P, _P, _P, _I, _I, _I, _OP, _OP, _I, _P],
    "skit_act_norm_bwd_reduce": [_P, _I, _I, _P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P],
    "skit_norm_bwd_apply_ex": [_P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _D, _P, _P, _P, _OP, _I, _P],
    "skit_mask_mul": [_P, _P, _I, _I, _I, _I, _P],
    "skit_channel_mean": [_P, _I, _I, _I, _I, _P, _P],
    "skit_channel_mean_bwd": [_P, _I, _I, _I, _I, _P, _P],
    "skit_norm_bwd_apply": [_P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _D, _P, _P, _OP, _I, _P],
    "skit_blur_down_fwd": [_P, _I, _I, _I, _I, _P, _P],
    "skit_blur_down_bwd": [_P, _I, _I, _I, _I, _P, _P],
    "skit_blur_up_fwd": [_P, _I, _I, _I, _I, _P, _P],
    "skit_blur_up_bwd": [_P, _I, _I, _I, _I, _P, _P],
    "skit_nchw_cat_to_operand": [C.POINTER(_P), C.POINTER(_I), _I, _I, _I, _I, _OP, _I, _I, _P],
    "skit_operand_grad_to_nchw": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P],
    "skit_g_head_fwd": [_P, _P, _I, _I, _I, _F, _P, _P, _P, _P],
    "skit_g_head_bwd": [_P, _P, _P, _P, _I, _I, _I, _OP, _I, _P],
    "skit_diffaug_bs_mask": [_P, _P, _P, _P, _I, _I, _I, _P, _P],
    "skit_avgpool3s2_fwd": [_P, _I, _I, _I, _P, _P],
    "skit_avgpool3s2_bwd": [_P, _I, _I, _I, _P, _I, _P],
    "skit_patch_gather": [C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), _I, _I, _I, _P, _P, _I, _I, _P, _I, _P],
    "skit_patch_scatter_add": [_P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P],
    "skit_gan_softplus": [_P, _I, _I, _F, _P, _P, _F, _P],
    "skit_gan_loss": [_P, _I, _I, _I, _I, _F, _P, _P, _F, _P],
    "skit_l1_loss": [_P, _P, _LL, _F, _P, _P, _F, _I, _P],
    "skit_adam_step": [_P, _P, _P, _P, _LL, _I, _F, _F, _F, _F, _F, _P],
    "skit_adam_step_dev": [_P, _P, _P, _P, _LL, _P, _F, _F, _F, _F, _P],
    "skit_debug_set_buffer": [_P],
    "skit_patch_sample_l2norm": [_P, _I, _I, _I, _P, _I, _P, _P, _P],
    "skit_patch_sample_l2norm_bwd": [_P, _P, _I, _I, _I, _P, _I, _P, _P],
    "skit_rows_scatter_add": [_P, _I, _I, _I, _P, _I, _P, _P],
    "skit_metric_minmax": [_P, _LL, _P, _P],
    "skit_metric_sq_err": [_P, _P, _LL, _P, _I, _P, _P],
    "skit_metric_ssim": [_P, _P, _I, _I, _I, _P, _F, _P, _P],
    "skit_metric_normal_angle": [_P, _P, _I, _I, _I, _F, _I, _P, _P],
    "skit_zero_bytes": [_P, _LL, _P],
    "skit_resize_u8": [_P, _I, _I, _I, _P, _I, _I, _I, _P],
    "skit_u8_crop_to_tensor": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "skit_contact_centers": [_P, _P, _P, _LL, _P, _P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "skit_touch_squares": [_P, _P, _P, _P, _P, _P, _I, _P, _I, _I, _P, _P, _I, _P, _P, _P, _I, _I, _P, _P, _P],
    "skit_laplacian_var_u8": [_P, _I, _I, _P, _P, _I, _I, _I, _P, _P],
    "skit_mask_box_bits": [_P, _I, _I, _I, _I, _P, _P, _P, _P],
    "skit_patchnce": [_P, _P, _I, _I, _I, _F, _P, _P, _F, _P],
}
_NO_RC = {"skit_last_error": (C.c_char_p, []), "skit_version": (_I, []), "skit_built_arch": (_I, [])}

_lib = None
launches = 0  # kernel-launching C-ABI calls made through call(); bench.py reports it


def load():
    """dlopen the in-tree library.  Raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libskit_b200.so is missing at %s — build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU/PyTorch fallback for the hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        if not hasattr(lib, name) and os.environ.get("SKIT_B200_LIB"):
            continue    # an older build loaded for an A/B measurement may lack the newest entry points
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _I
    for name, (res, argtypes) in _NO_RC.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = res
    _lib = lib
    return lib


def call(name, *args):
    global launches
    lib = load()
    rc = getattr(lib, name)(*args)
    launches += 1
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.skit_last_error().decode()))


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "skit_b200 ops need contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())
