"""TEST INFRASTRUCTURE ONLY (oracle/): golden vectors for the data pipeline.

Builds a small seeded synthetic dataset in the reference's on-disk format (`trainS/ trainI/ trainM/ trainT/ valT/`,
`*_tactile.npz` with the keys `data/dataset_util.py:18` documents), runs the UNMODIFIED reference
`data/singleskit_dataset.py:SingleSkitDataset` on it (imported in place through oracle/ref_loader.py, CPU) and stores
what its `data_dict` holds, plus PIL resize outputs, under tests/golden/data_pipeline.npz.

    python oracle/make_data_golden.py            # needs /root/reference; run in the build container only
"""
import argparse
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_dataset(root, seed=0, W=300, H=290, n_train=14, n_val=6, patch_h=(52, 72), patch_w=(56, 80), patch_x=(72, 112), patch_y=(72, 118),
                  strokes=40, pad_offset=(0, 0)):
    """Write the synthetic dataset; everything derives from `seed` (numpy Generator), PNG is lossless."""
    from PIL import Image
    g = np.random.default_rng(seed)
    for d in ("trainS", "trainI", "trainM", "trainT", "valT"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    yy, xx = np.mgrid[0:H, 0:W]
    # sketch: white background with dark strokes
    S = np.full((H, W), 255, np.uint8)
    for _ in range(strokes):
        x0, y0 = g.integers(0, W), g.integers(0, H)
        ang = g.uniform(0, np.pi)
        t = np.arange(0, g.integers(30, 160))
        px = np.clip((x0 + t * np.cos(ang)).astype(int), 0, W - 1)
        py = np.clip((y0 + t * np.sin(ang)).astype(int), 0, H - 1)
        S[py, px] = g.integers(0, 90)
        S[np.clip(py + 1, 0, H - 1), px] = g.integers(0, 90)
    I = (127 + 90 * np.sin(xx[..., None] / (7.0 + np.arange(3)) + yy[..., None] / 11.0) + g.normal(0, 3, (H, W, 3))).clip(0, 255).astype(np.uint8)
    M = ((((xx - W / 2) / (0.42 * W)) ** 2 + ((yy - H / 2) / (0.40 * H)) ** 2) <= 1).astype(np.uint8) * 255
    Image.fromarray(S, "L").save(os.path.join(root, "trainS", "syn.png"))
    Image.fromarray(I, "RGB").save(os.path.join(root, "trainI", "syn.png"))
    Image.fromarray(M, "L").save(os.path.join(root, "trainM", "syn.png"))
    for sub, n in (("trainT", n_train), ("valT", n_val)):
        for i in range(n):
            h, w = int(g.integers(*patch_h)), int(g.integers(*patch_w))
            # every patch inside the region every admissible crop covers: the reference indexes its valid-patch lists by the raw
            # patch index (`singleskit_dataset.py:742-743`), which only works when no patch is rejected
            x, y = int(g.integers(*patch_x)) - pad_offset[0], int(g.integers(*patch_y)) - pad_offset[1]
            gx = g.uniform(-0.3, 0.3, (h, w)).astype(np.float32)
            gy = g.uniform(-0.3, 0.3, (h, w)).astype(np.float32)
            ty, tx = np.mgrid[0:h, 0:w]
            cy, cx = h / 2 + g.uniform(-4, 4), w / 2 + g.uniform(-4, 4)
            blob = (((tx - cx) / (0.45 * w)) ** 2 + ((ty - cy) / (0.45 * h)) ** 2) <= 1
            core = (((tx - cx) / (0.12 * w)) ** 2 + ((ty - cy) / (0.12 * h)) ** 2) <= 1
            core &= (ty >= 16) & (ty <= h - 16) & (tx >= 16) & (tx <= w - 16)
            # one file per sub-directory: the reference lists directories sorted but file names in OS order (image_folder.py:52-57)
            os.makedirs(os.path.join(root, sub, "p%02d" % i), exist_ok=True)
            np.savez(os.path.join(root, sub, "p%02d" % i, "syn_%02d_tactile.npz" % i), gx_raw=gx, gy_raw=gy,
                     vision_mask_x=x, vision_mask_y=y, vision_mask_h=h, vision_mask_w=w,
                     touch_thresh=blob.astype(np.uint8) * 255, touch_center_thresh=core.astype(np.uint8) * 255)
    return root


def dataset_options(root, **kw):
    """The option fields `SingleSkitDataset` reads (`data/singleskit_dataset.py:84-196`, `models/sinskitG_model.py:202-300`)."""
    o = dict(dataroot=root, subdir_S="trainS", subdir_I="trainI", subdir_T="trainT", subdir_M="trainM", subdir_valT="valT",
             is_train=True, isTrain=True, max_dataset_size=float("inf"), sketch_nc=1, image_nc=3, use_bg_mask=True,
             preprocess="crop", random_scale_max=3.0, batch_size=1, crop_size=256, center_w=160, center_h=150, data_len=4,
             w_resampling=True, resampling_w_min=1, resampling_w_max=10, T_resolution_multiplier=1, batch_size_G2=8,
             batch_size_G2_val=16, sample_bbox_per_patch=2, serial_batches=False, num_threads=0)
    o.update(kw)
    return argparse.Namespace(**o)


CASES = {
    "crop": dict(data_len=3),                                                                   # the reference's training default: crop only
    "zoom_crop": dict(preprocess="zoom_crop", random_scale_max=1.5, crop_size=192, data_len=2),   # LANCZOS zoom, then 192 -> 256 "power 2" resize
    "noresample": dict(w_resampling=False, sample_bbox_per_patch=1, data_len=1),
    # the test phase (models/sinskitG_model.py:359-374): no augmentation, centre crop, every patch, the centre-most squares, no validation set
    "test": dict(is_train=False, isTrain=False, preprocess="none", data_len=1, sample_bbox_per_patch=1, batch_size_G2=100, subdir_valT=None),
}


def run_reference(root, opt, seed):
    from oracle import ref_loader
    ref_loader.load_reference()
    cwd = os.getcwd()
    os.chdir(root)      # the reference creates ./logs/<date> (myutils.py:14-29)
    try:
        import contextlib
        import io
        from data.singleskit_dataset import SingleSkitDataset
        random.seed(seed); np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            ds = SingleSkitDataset(opt)
        return ds
    finally:
        os.chdir(cwd)


SKIT_MATERIALS = ("matA", "matB")
SKIT_CASE = dict(preprocess="zoom_crop", random_scale_max=1.3, crop_size=256, data_len=3, material_list=list(SKIT_MATERIALS), padded_size=300,
                 load_contact_mask=True)


def synth_skit_datasets(base):
    """Two materials in the directory layout `skit_dataset.py:141` expects under <base>/datasets/.  The ROIs are stored in the unpadded
    (center_w x center_h) frame, as the reference's "padded" data does (`global_padding_find_coords`, dataset_util.py:240-243)."""
    for k, mat in enumerate(SKIT_MATERIALS):
        synth_dataset(os.path.join(base, "datasets", "singleskit_%s_padded_300_x1" % mat), seed=10 + k, pad_offset=((300 - 160) // 2, (300 - 150) // 2))
    return base


def synth_skit_external(base):
    """The `_edit0` directories `use_external_test_input` reads (skit_dataset.py:116-137): sketch + mask of one material, image + mask
    of the style material."""
    import shutil
    synth_skit_datasets(base)
    for mat in SKIT_MATERIALS:
        src = os.path.join(base, "datasets", "singleskit_%s_padded_300_x1" % mat)
        dst = src + "_edit0"
        if not os.path.exists(dst):
            shutil.copytree(src, dst)
    return base


def skit_external_options(base):
    return dataset_options(os.path.join("./datasets", "singleskit_%s_padded_300_x1/" % SKIT_MATERIALS[0]), is_train=False, isTrain=False,
                           preprocess="none", data_len=1, material_list=[SKIT_MATERIALS[0]], padded_size=300, load_contact_mask=True,
                           use_external_test_input=True, test_sketch_material=SKIT_MATERIALS[0], test_style_material=SKIT_MATERIALS[1],
                           subdir_valT=None)


def skit_options(base):
    return dataset_options(os.path.join("./datasets", "singleskit_%s_padded_300_x1/" % SKIT_MATERIALS[0]), **SKIT_CASE)


def run_reference_skit(base, opt, seed):
    from oracle import ref_loader
    ref_loader.load_reference()
    cwd = os.getcwd()
    os.chdir(base)      # material directories are relative to the working directory (skit_dataset.py:141)
    try:
        import contextlib
        import io
        from data.skit_dataset import SkitDataset
        random.seed(seed); np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            ds = SkitDataset(opt)
        return ds
    finally:
        os.chdir(cwd)


def flatten(prefix, item, out):
    import torch
    from oracle import data_oracle as DO
    for k, v in item.items():
        key = prefix + "/" + k
        if k in ("S", "I", "M"):      # full-resolution tensors go in as the bytes they were made from (checked lossless)
            t = v.numpy()
            u8 = np.round((t * 0.5 + 0.5) * 255 if k != "M" else t * 255).astype(np.uint8).transpose(1, 2, 0)
            assert np.array_equal(DO.to_tensor_norm(u8, normalize=k != "M"), t), k
            out[key + "_u8"] = u8
        elif isinstance(v, torch.Tensor):
            out[key] = v.numpy()
        elif isinstance(v, np.ndarray):
            out[key] = v
        elif isinstance(v, dict):
            out[key] = np.array([float(v[a]) for a in sorted(v)], np.float64)
            out[key + "__keys"] = np.array(sorted(v))
        elif isinstance(v, (list, tuple)) and len(v) and not isinstance(v[0], str):
            out[key] = np.array(v)
        elif isinstance(v, (list, tuple)) and len(v) == 0:
            out[key] = np.zeros((0,))


def main():
    root = "/tmp/vts_data_golden"
    synth_dataset(root)
    out = {}
    for name, kw in CASES.items():
        opt = dataset_options(root, **kw)
        ds = run_reference(root, opt, seed=123)
        for idx in range(len(ds)):
            flatten("%s/%d" % (name, idx), ds[idx], out)
        out[name + "/len"] = np.array(len(ds))
    base = synth_skit_datasets("/tmp/vts_data_golden_skit")
    ds = run_reference_skit(base, skit_options(base), seed=321)
    for idx in range(len(ds)):
        flatten("skit/%d" % idx, ds[idx], out)
        out["skit/%d/name" % idx] = np.array(ds[idx]["name"])
        out["skit/%d/M_paths" % idx] = np.array(ds[idx]["M_paths"])
    out["skit/len"] = np.array(len(ds))
    base = synth_skit_external("/tmp/vts_data_golden_skit")
    ds = run_reference_skit(base, skit_external_options(base), seed=11)
    item = ds[0]
    flatten("skit_ext/0", {k: v for k, v in item.items() if k not in ("style_I", "style_M")}, out)
    for k in ("style_I", "style_M"):
        t = item[k].numpy()
        out["skit_ext/0/%s_u8" % k] = np.round((t * 0.5 + 0.5) * 255 if k == "style_I" else t * 255).astype(np.uint8).transpose(1, 2, 0)
    out["skit_ext/0/keys"] = np.array(sorted(item))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "data_pipeline.npz"), **out)
    print("wrote", len(out), "arrays")
    for k in sorted(out)[:40]:
        print(k, out[k].shape, out[k].dtype)


if __name__ == "__main__":
    main()
