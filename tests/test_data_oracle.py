"""CPU: the data-pipeline oracle (oracle/data_oracle.py) pinned against Pillow / OpenCV themselves and against what the
reference's own SingleSkitDataset produced (tests/golden/data_pipeline.npz, made by oracle/make_data_golden.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import data_oracle as DO  # noqa: E402
from oracle import make_data_golden as MG  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "data_pipeline.npz")


@pytest.mark.parametrize("shape", [(40, 50, 1, 27, 33), (40, 50, 3, 61, 77), (300, 290, 3, 270, 221), (33, 47, 1, 33, 20),
                                   (64, 64, 3, 256, 256), (101, 77, 3, 13, 9), (17, 19, 1, 1, 1)])
def test_resize_oracle_matches_pillow(shape):
    Image = pytest.importorskip("PIL.Image")
    h, w, c, oh, ow = shape
    g = np.random.default_rng(h * w + c)
    a = g.integers(0, 256, (h, w, c), dtype=np.uint8)
    a = a[:, :, 0] if c == 1 else a
    for kind, pk in ((DO.LANCZOS, Image.LANCZOS), (DO.BICUBIC, Image.BICUBIC), (DO.BILINEAR, Image.BILINEAR), (DO.BOX, Image.BOX),
                     (DO.HAMMING, Image.HAMMING)):
        ref = np.array(Image.fromarray(a).resize((ow, oh), pk))
        assert np.array_equal(ref, DO.pil_resize_u8(a, oh, ow, kind)), (shape, kind)      # bit-exact


def test_laplacian_oracle_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    g = np.random.default_rng(3)
    for size in (32, 31, 8):
        p = g.integers(0, 256, (size, size), dtype=np.uint8)
        p[p > 100] = 255
        ref = cv2.Laplacian(p - np.ones_like(p) * 255, cv2.CV_64F).var()       # util/util.py:261-265 with ref = white
        assert DO.laplacian_var_u8(p) == pytest.approx(ref, rel=1e-13)


def test_to_tensor_oracle_matches_torchvision():
    tv = pytest.importorskip("torchvision.transforms")
    a = np.random.default_rng(0).integers(0, 256, (9, 7, 3), dtype=np.uint8)
    ref = tv.Compose([tv.ToTensor(), tv.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])(a).numpy()
    assert np.array_equal(ref, DO.to_tensor_norm(a))


def _load_sources(root):
    from PIL import Image
    S = np.array(Image.open(os.path.join(root, "trainS", "syn.png")).convert("L"))
    I = np.array(Image.open(os.path.join(root, "trainI", "syn.png")).convert("RGB"))
    M = np.array(Image.open(os.path.join(root, "trainM", "syn.png")).convert("L"))
    return S, I, M


def _final_u8(img, ap, crop):
    """zoom -> ratio resize -> crop -> power-of-2 resize with the oracle, from the golden augmentation parameters."""
    h, w = img.shape[:2]
    if ap["scale_factor_h"] != 1 or ap["scale_factor_w"] != 1:
        img = DO.pil_resize_u8(img, int(round(h * ap["scale_factor_h"])), int(round(w * ap["scale_factor_w"])))
    h, w = img.shape[:2]
    img = DO.pil_resize_u8(img, int(round(h * ap["resize_ratio"])), int(round(w * ap["resize_ratio"])))
    x, y = int(ap["crop_pos_x"]), int(ap["crop_pos_y"])
    img = img[y:y + crop, x:x + crop]
    t = int(round(crop / 256) * 256)
    return DO.pil_resize_u8(img, t, t)


@pytest.mark.parametrize("case", ["crop", "zoom_crop"])
def test_oracle_reproduces_reference_dataset_images(case, tmp_path):
    pytest.importorskip("PIL.Image")
    d = np.load(GOLD)
    root = MG.synth_dataset(str(tmp_path / "ds"))
    S, I, M = _load_sources(root)
    crop = MG.CASES[case].get("crop_size", 256)
    for idx in range(int(d[case + "/len"])):
        ap = dict(zip(d["%s/%d/augmentation_params__keys" % (case, idx)], d["%s/%d/augmentation_params" % (case, idx)]))
        assert np.array_equal(_final_u8(S, ap, crop), d["%s/%d/S_u8" % (case, idx)][:, :, 0])
        assert np.array_equal(_final_u8(I, ap, crop), d["%s/%d/I_u8" % (case, idx)])
        assert np.array_equal(_final_u8(M, ap, crop), d["%s/%d/M_u8" % (case, idx)][:, :, 0])


def test_oracle_reproduces_reference_touch_squares(tmp_path):
    """Every square the reference kept (T_coords row) is one of the oracle's valid contact centres of that patch, and its I_masks
    entry is the oracle's touch_mask x M_patch / 255 window."""
    pytest.importorskip("PIL.Image")
    d = np.load(GOLD)
    root = MG.synth_dataset(str(tmp_path / "ds"))
    M3 = d["crop/0/M_u8"][:, :, 0]
    files = sorted(os.path.join(r, f) for r, _, fs in os.walk(os.path.join(root, "trainT")) for f in fs)
    rois = {}
    for f in files:
        z = np.load(f)
        rois[(int(z["vision_mask_h"]), int(z["vision_mask_w"]))] = z
    coords, masks = d["crop/0/T_coords"], d["crop/0/I_masks"]
    for row, mask in zip(coords, masks):
        z = rois[(int(row[2]), int(row[3]))]
        tm, cm = z["touch_thresh"] / 255, z["touch_center_thresh"] / 255
        centres = DO.contact_centers(tm, cm, M3, int(row[0]), int(row[1]))
        cx, cy = int(row[6]) + 16, int(row[7]) + 16
        assert (cx, cy) in centres
        win = tm[cy - 16:cy + 16, cx - 16:cx + 16] * DO.crop_zero(M3, int(row[0]) + cx - 16, int(row[1]) + cy - 16, 32) / 255
        assert np.array_equal(win, mask)


@pytest.mark.parametrize("case", ["crop", "zoom_crop", "noresample", "test"])
def test_dataset_oracle_reproduces_reference_items(case, tmp_path):
    """The numpy restatement of the whole dataset (oracle `dataset_items`), seeded like the golden run, gives the reference's items."""
    import random
    pytest.importorskip("PIL.Image")
    d = np.load(GOLD)
    root = MG.synth_dataset(str(tmp_path / "ds"))
    opt = MG.dataset_options(root, **MG.CASES[case])
    random.seed(123); np.random.seed(123)
    items = DO.dataset_items(opt)
    assert len(items) == int(d[case + "/len"])
    for idx, item in enumerate(items):
        pre = "%s/%d/" % (case, idx)
        assert np.array_equal(item["S_u8"], d[pre + "S_u8"][:, :, 0]) and np.array_equal(item["I_u8"], d[pre + "I_u8"])
        assert np.array_equal(item["M_u8"], d[pre + "M_u8"][:, :, 0])
        for k in ("T_images", "T_coords", "I_masks", "full_T_coords", "val_T_images", "val_T_coords", "val_I_masks", "val_full_T_coords"):
            if k.startswith("val_") and d[pre + k].size == 0:       # the test phase has no validation set: the reference stores []
                assert k not in item
                continue
            assert np.array_equal(item[k], d[pre + k]), (case, idx, k)
        keys = list(d[pre + "augmentation_params__keys"])
        assert np.array_equal(np.array([float(item["augmentation_params"][a]) for a in keys]), d[pre + "augmentation_params"])


def test_resize_oracle_random_shapes_against_pillow():
    """Randomised sweep (seeded): up- and down-scaling by arbitrary ratios, 1 and 3 bands, extreme aspect ratios — bit-exact with Pillow."""
    Image = pytest.importorskip("PIL.Image")
    g = np.random.default_rng(2024)
    for _ in range(24):
        h, w = int(g.integers(1, 90)), int(g.integers(1, 90))
        oh, ow = int(g.integers(1, 120)), int(g.integers(1, 120))
        c = int(g.choice([1, 3]))
        a = g.integers(0, 256, (h, w, c), dtype=np.uint8)
        a = a[:, :, 0] if c == 1 else a
        for kind, pk in ((DO.LANCZOS, Image.LANCZOS), (DO.BICUBIC, Image.BICUBIC)):
            ref = np.array(Image.fromarray(a).resize((ow, oh), pk))
            assert np.array_equal(ref, DO.pil_resize_u8(a, oh, ow, kind)), (h, w, c, oh, ow, kind)
