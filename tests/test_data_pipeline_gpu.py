"""GPU: the data-pipeline kernels (csrc/data_ops.cu, through the C ABI) against the oracle (oracle/data_oracle.py, itself pinned to
Pillow / OpenCV by tests/test_data_oracle.py), and the device `SingleSkitDataset` against the items the reference's own dataset
class produced from the same seeded synthetic directory (tests/golden/data_pipeline.npz).  Integer / byte work: bit-exact."""
import argparse
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import data_oracle as DO  # noqa: E402
from oracle import make_data_golden as MG  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden", "data_pipeline.npz")


@pytest.fixture(scope="module")
def V():
    import vts_b200
    return vts_b200


def _dev(a):
    a = a[:, :, None] if a.ndim == 2 else a
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape", [(40, 50, 1, 27, 33), (40, 50, 3, 61, 77), (300, 290, 3, 270, 221), (33, 47, 1, 33, 20), (64, 64, 3, 256, 256),
                                   (101, 77, 3, 13, 9), (17, 19, 1, 1, 1), (1800, 1800, 3, 1371, 1618), (192, 192, 1, 256, 256)])
def test_resize_matches_oracle(V, shape):
    h, w, c, oh, ow = shape
    a = np.random.default_rng(h + w + c).integers(0, 256, (h, w, c), dtype=np.uint8)
    kinds = (DO.LANCZOS, DO.BICUBIC, DO.BILINEAR, DO.BOX, DO.HAMMING) if h < 1000 else (DO.LANCZOS,)
    for kind in kinds:
        out = V.data_pipeline.resize_u8(_dev(a), oh, ow, kind).cpu().numpy()
        assert np.array_equal(out, DO.pil_resize_u8(a, oh, ow, kind)), (shape, kind)


def test_resize_same_size_is_a_copy_and_bad_filter_raises(V):
    a = np.random.default_rng(0).integers(0, 256, (20, 30, 3), dtype=np.uint8)
    assert np.array_equal(V.data_pipeline.resize_u8(_dev(a), 20, 30).cpu().numpy(), a)
    with pytest.raises(RuntimeError, match="unsupported filter"):
        V.data_pipeline.resize_u8(_dev(a), 10, 10, 0)       # NEAREST is not a convolution filter in Pillow either
    with pytest.raises(RuntimeError, match="CUDA uint8"):
        V.data_pipeline.resize_u8(torch.zeros(4, 4, 1, dtype=torch.uint8), 2, 2)


@pytest.mark.parametrize("box", [(3, 5, 20, 17), (-4, -2, 16, 16), (30, 20, 24, 24)])
def test_crop_to_tensor_matches_oracle(V, box):
    x0, y0, w, h = box
    a = np.random.default_rng(1).integers(0, 256, (37, 41, 3), dtype=np.uint8)
    for normalize in (True, False):
        out = V.data_pipeline.crop_to_tensor(_dev(a), x0, y0, w, h, normalize).cpu().numpy()
        ref = np.zeros((h, w, 3), np.uint8)
        ys, xs, ye, xe = max(0, y0), max(0, x0), min(37, y0 + h), min(41, x0 + w)
        ref[ys - y0:ye - y0, xs - x0:xe - x0] = a[ys:ye, xs:xe]
        assert np.array_equal(out, DO.to_tensor_norm(ref, normalize))


def _touch_files(tmp, n=5, seed=4):
    g = np.random.default_rng(seed)
    paths = []
    for i in range(n):
        h, w = int(g.integers(40, 70)), int(g.integers(40, 90))
        tm = (g.random((h, w)) < 0.02).astype(np.uint8) * 255          # sparse contact pixels: windows with and without a hit
        cm = np.zeros((h, w), np.uint8)
        cm[16:h - 15, 16:w - 15] = (g.random((h - 31, w - 31)) < 0.5) * 255
        p = os.path.join(tmp, "t%d_tactile.npz" % i)
        np.savez(p, gx_raw=g.normal(size=(h, w)).astype(np.float32), gy_raw=g.normal(size=(h, w)).astype(np.float32),
                 vision_mask_x=int(g.integers(-10, 60)), vision_mask_y=int(g.integers(-10, 60)), vision_mask_h=h, vision_mask_w=w,
                 touch_thresh=tm, touch_center_thresh=cm)
        paths.append(p)
    return paths


def test_contact_centres_squares_and_weights_match_oracle(V, tmp_path):
    paths = _touch_files(str(tmp_path))
    g = np.random.default_rng(9)
    M3 = ((g.random((120, 130)) < 0.6) * 255).astype(np.uint8)
    M3[:30] = 0
    S3 = g.integers(0, 256, (120, 130), dtype=np.uint8)
    ts = V.data_pipeline.TouchSet(paths, torch.device("cuda"))
    roi_x = [int(r[0]) for r in ts.roi]
    roi_y = [int(r[1]) for r in ts.roi]
    counts, in_mask = ts.contact_centers(_dev(M3), roi_x, roi_y)
    sel_patch, sel_rank, want = [], [], []
    for i, p in enumerate(paths):
        z = np.load(p)
        tm, cm = z["touch_thresh"] / 255, z["touch_center_thresh"] / 255
        ref = DO.contact_centers(tm, cm, M3, roi_x[i], roi_y[i])
        assert counts[i] == len(ref), i
        h, w = tm.shape
        rect = DO.crop_zero(M3, roi_x[i], roi_y[i], max(h, w))[:h, :w]
        assert bool(in_mask[i]) == bool(rect.any())
        lin = ts.centers[int(ts.pix_off_host[i]):int(ts.pix_off_host[i]) + len(ref)].cpu().numpy()
        assert [(int(v % w), int(v // w)) for v in lin] == ref               # same pixels, same (np.where) order
        for k in (0, len(ref) // 2, len(ref) - 1):
            if len(ref):
                sel_patch.append(i); sel_rank.append(k); want.append((i, ref[k]))
    cx, cy, T, Mk = ts.squares(_dev(M3), sel_patch, sel_rank)
    xs, ys = [], []
    for k, (i, (rx, ry)) in enumerate(want):
        assert (int(cx[k]), int(cy[k])) == (rx, ry)
        z = np.load(paths[i])
        win = (slice(ry - 16, ry + 16), slice(rx - 16, rx + 16))
        assert np.array_equal(T[k, 0].cpu().numpy(), z["gx_raw"][win]) and np.array_equal(T[k, 1].cpu().numpy(), z["gy_raw"][win])
        mp = DO.crop_zero(M3, roi_x[i] + rx - 16, roi_y[i] + ry - 16, 32)
        assert np.array_equal(Mk[k].cpu().numpy(), (z["touch_thresh"] / 255)[win] * mp / 255)
        xs.append(roi_x[i] + rx - 16); ys.append(roi_y[i] + ry - 16)
    var = V.data_pipeline.laplacian_var(_dev(S3), xs, ys, 32)
    for k in range(len(xs)):
        assert var[k] == pytest.approx(DO.laplacian_var_u8(DO.crop_zero(S3, xs[k], ys[k], 32)), rel=1e-12)      # fp64: tolerance 1e-12


def test_touch_set_rejects_centres_near_the_border(V, tmp_path):
    cm = np.zeros((40, 40), np.uint8); cm[5, 20] = 255
    p = str(tmp_path / "b_tactile.npz")
    np.savez(p, gx_raw=np.zeros((40, 40), np.float32), gy_raw=np.zeros((40, 40), np.float32), vision_mask_x=0, vision_mask_y=0,
             vision_mask_h=40, vision_mask_w=40, touch_thresh=np.ones((40, 40), np.uint8), touch_center_thresh=cm)
    with pytest.raises(ValueError, match="within 16 pixels"):
        V.data_pipeline.TouchSet([p], torch.device("cuda"))


def _check_item(d, prefix, item):
    for k in ("S", "I", "M"):
        ref = DO.to_tensor_norm(d["%s/%s_u8" % (prefix, k)], normalize=k != "M")
        assert item[k].is_cuda and np.array_equal(item[k].cpu().numpy(), ref), (prefix, k)
    for k in ("T_images", "I_masks", "val_T_images", "val_I_masks"):
        ref = d["%s/%s" % (prefix, k)]
        if ref.size == 0:                       # the test phase has no validation set: the reference stores []
            assert len(item[k]) == 0
            continue
        got = item[k].cpu().numpy()
        assert got.dtype == ref.dtype and np.array_equal(got, ref), (prefix, k)
    for k in ("T_coords", "val_T_coords", "full_T_coords", "val_full_T_coords"):
        assert np.array_equal(np.asarray(item[k]).reshape(d["%s/%s" % (prefix, k)].shape), d["%s/%s" % (prefix, k)]), (prefix, k)
    keys = list(d[prefix + "/augmentation_params__keys"])
    assert sorted(item["augmentation_params"]) == keys
    assert np.array_equal(np.array([float(item["augmentation_params"][a]) for a in keys]), d[prefix + "/augmentation_params"])


@pytest.mark.parametrize("case", ["crop", "zoom_crop", "noresample", "test"])
def test_dataset_matches_reference_golden(V, case, tmp_path):
    """Same seeded directory, same seeds, same options -> the reference's items bit for bit (LANCZOS zoom and power-of-2 resize
    included in the zoom_crop case)."""
    d = np.load(GOLD)
    root = MG.synth_dataset(str(tmp_path / "ds"))
    opt = MG.dataset_options(root, **MG.CASES[case])
    random.seed(123); np.random.seed(123)
    ds = V.SingleSkitDataset(opt)
    assert len(ds) == int(d[case + "/len"])
    for idx in range(len(ds)):
        _check_item(d, "%s/%d" % (case, idx), ds[idx])
    assert ds[0]["S"] is ds[0]["S"]            # cached like the reference's data_dict entries


def test_dataset_feeds_the_train_step(V, tmp_path):
    """Items go through a torch DataLoader (the reference's CustomDatasetDataLoader settings) into set_input / optimize_parameters."""
    root = MG.synth_dataset(str(tmp_path / "ds"))
    opt = MG.dataset_options(root, data_len=2)
    random.seed(1); np.random.seed(1)
    ds = V.SingleSkitDataset(opt)
    loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=0, drop_last=True)
    mopt = V.default_options(netG="unet256_custom", ngf=4, ndf=4, crop_size=256, batch_size_G2=8, add_fake_T_sample_size=4)
    m = V.SinSKITGModel(mopt)
    before = ds[0]["S"].clone()
    for _ in range(2):
        for batch in loader:
            m.set_input(batch)
            m.optimize_parameters()
    assert m.h2d_bytes < 4096                          # nothing but the patch offsets crosses the bus
    ref_t = V.model_utils.random_patch_offset_table(ds[1]["M"][None].cpu())        # host construction on the item's mask
    got_t = V.model_utils.offset_table_from_arrays(ds[1]["M_box_bits"].numpy(), ds[1]["M_box_rowcount"].numpy(), 256, 256)
    assert len(ref_t) == len(got_t) and np.array_equal(ref_t.rows, got_t.rows) and np.array_equal(ref_t.cols, got_t.cols)
    assert torch.equal(ds[0]["S"], before)             # the model masks its own staged copy, not the dataset's cached tensor
    losses = m.get_current_losses()
    assert all(np.isfinite(float(v)) for v in losses.values())


def test_dataset_rejects_patches_outside_the_crop(V, tmp_path):
    root = MG.synth_dataset(str(tmp_path / "ds"))
    z = dict(np.load(os.path.join(root, "trainT", "p00", "syn_00_tactile.npz")))
    z["vision_mask_x"] = 0; z["vision_mask_y"] = 0
    np.savez(os.path.join(root, "trainT", "p00", "syn_00_tactile.npz"), **z)
    opt = MG.dataset_options(root, center_w=0, center_h=0, data_len=8)
    random.seed(5); np.random.seed(5)
    with pytest.raises(IndexError, match="valid-patch bookkeeping"):
        V.SingleSkitDataset(opt)


def test_host_items_go_through_a_pinning_loader(V, tmp_path):
    """The reference's loader settings (data/__init__.py:75-82: pin_memory=True) with `host_items`: same values, host tensors."""
    d = np.load(GOLD)
    root = MG.synth_dataset(str(tmp_path / "ds"))
    opt = MG.dataset_options(root, **MG.CASES["crop"])
    random.seed(123); np.random.seed(123)
    ds = V.SingleSkitDataset(opt, host_items=True)
    loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=0, drop_last=True, pin_memory=True)
    for idx, batch in enumerate(loader):
        assert not batch["S"].is_cuda and batch["S"].is_pinned()
        item = {k: (v[0] if torch.is_tensor(v) else v) for k, v in batch.items()}
        ref = DO.to_tensor_norm(d["crop/%d/S_u8" % idx])
        assert np.array_equal(item["S"].numpy(), ref)
        assert np.array_equal(item["T_images"].numpy(), d["crop/%d/T_images" % idx])
        assert np.array_equal(item["T_coords"].numpy(), d["crop/%d/T_coords" % idx])


def test_multi_material_dataset_matches_reference_golden(V, tmp_path, monkeypatch):
    """`SkitDataset` (data/skit_dataset.py): two materials round-robin, a LANCZOS zoom level per index, ROIs through the "padded"
    offset, the ratio resize of a zoomed source smaller than the crop — against the reference class's own items."""
    d = np.load(GOLD)
    base = MG.synth_skit_datasets(str(tmp_path / "skit"))
    monkeypatch.chdir(base)                       # material directories are relative to the working directory, as in the reference
    opt = MG.skit_options(base)
    random.seed(321); np.random.seed(321)
    ds = V.SkitDataset(opt)
    assert len(ds) == int(d["skit/len"])
    for idx in range(len(ds)):
        item = ds[idx]
        _check_item(d, "skit/%d" % idx, item)
        assert item["name"] == str(d["skit/%d/name" % idx]) and item["M_paths"] == str(d["skit/%d/M_paths" % idx])
    opt2 = MG.skit_options(base)
    opt2.load_contact_mask = False
    with pytest.raises(NotImplementedError, match="load_contact_mask"):
        V.SkitDataset(opt2)


@pytest.mark.parametrize("hw", [(64, 80), (97, 131), (300, 290)])
def test_device_offset_table_matches_host_table(V, hw):
    """The random-patch candidate table built by the kernel (bit map + row counts) lists the same (row, col) positions, in the same
    order, as the host table (whose construction tests/test_host_logic.py pins against the reference's conv2d + nonzero)."""
    import random as pyrandom
    MU = V.model_utils
    h, w = hw
    g = np.random.default_rng(h)
    for kind in ("ellipse", "sparse", "ones", "zeros"):
        if kind == "ellipse":
            yy, xx = np.mgrid[0:h, 0:w]
            m = ((((xx - w / 2) / (0.3 * w)) ** 2 + ((yy - h / 2) / (0.35 * h)) ** 2) <= 1).astype(np.float32)
        elif kind == "sparse":
            m = (g.random((h, w)) < 0.002).astype(np.float32)
        else:
            m = np.full((h, w), 1.0 if kind == "ones" else 0.0, np.float32)
        M = torch.from_numpy(m)[None, None]
        host, dev = MU.random_patch_offset_table(M), MU.random_patch_offset_table(M.cuda())
        assert len(host) == len(dev), kind
        assert np.array_equal(host.rows, dev.rows) and np.array_equal(host.cols, dev.cols), kind
        if len(host) >= 8:
            a = host.sample(8, rng=pyrandom.Random(3)); b = dev.sample(8, rng=pyrandom.Random(3))
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_float64_touch_maps_and_sketch_only_dataset(V, tmp_path):
    """Touch maps stored as float64 keep their dtype through the gather (ToTensor on a float64 array gives a DoubleTensor in the
    reference too); a directory without trainI ("edit" sketches, singleskit_dataset.py:139-145) yields sketch / mask items only."""
    import shutil
    root = MG.synth_dataset(str(tmp_path / "ds"))
    for sub in ("trainT", "valT"):
        for r, _, fs in os.walk(os.path.join(root, sub)):
            for f in fs:
                z = dict(np.load(os.path.join(r, f)))
                z["gx_raw"] = z["gx_raw"].astype(np.float64); z["gy_raw"] = z["gy_raw"].astype(np.float64)
                np.savez(os.path.join(r, f), **z)
    opt = MG.dataset_options(root, data_len=1)
    random.seed(2); np.random.seed(2)
    item = V.SingleSkitDataset(opt)[0]
    assert item["T_images"].dtype == torch.float64 and item["T_images"].shape[1:] == (2, 32, 32)
    # the same draw on float32 copies gives the same values
    ref_root = MG.synth_dataset(str(tmp_path / "ds32"))
    random.seed(2); np.random.seed(2)
    ref = V.SingleSkitDataset(MG.dataset_options(ref_root, data_len=1))[0]
    assert torch.equal(item["T_images"].float(), ref["T_images"]) and np.array_equal(item["T_coords"], ref["T_coords"])
    # sketch-only ("edit") directory
    edit = str(tmp_path / "ds_edit0")
    shutil.copytree(root, edit)
    shutil.rmtree(os.path.join(edit, "trainI"))
    random.seed(2); np.random.seed(2)
    e = V.SingleSkitDataset(MG.dataset_options(edit, data_len=2))[1]
    assert set(e) >= {"S", "M", "name", "S_paths", "M_paths", "augmentation_params"} and "I" not in e and e["T_images"] == []
    assert tuple(e["S"].shape) == (1, 256, 256) and tuple(e["M"].shape) == (1, 256, 256)


def test_multi_material_external_sketch_matches_reference_golden(V, tmp_path, monkeypatch):
    """`SkitDataset` with `use_external_test_input` (skit_dataset.py:116-137, test.py's path for new sketches): the sketch / mask of one
    `_edit0` directory and the style image / mask of another ride through the same centre crop — against the reference class's item."""
    d = np.load(GOLD)
    base = MG.synth_skit_external(str(tmp_path / "skit"))
    monkeypatch.chdir(base)
    random.seed(11); np.random.seed(11)
    item = V.SkitDataset(MG.skit_external_options(base))[0]
    assert sorted(k for k in item if not k.startswith("M_box")) == list(d["skit_ext/0/keys"])
    for k in ("S", "M", "style_I", "style_M"):
        ref = DO.to_tensor_norm(d["skit_ext/0/%s_u8" % k], normalize=not k.endswith("M"))
        assert np.array_equal(item[k].cpu().numpy(), ref), k
    assert item["T_images"] == []
    keys = list(d["skit_ext/0/augmentation_params__keys"])
    assert np.array_equal(np.array([float(item["augmentation_params"][a]) for a in keys]), d["skit_ext/0/augmentation_params"])
