"""Synthetic scenes with the shapes and dtypes of the reference's datasets.

No dataset ships with the reference and there is no network, so benchmarks and tests run on seeded random
scenes: 8-bit noise scaled to [0,1] like ``img_as_float`` (isprs:194-198) and block-constant label maps.
"""
import numpy as np

SHAPES = {
    "vaihingen": dict(H=2000, W=2500, C=4, K=6, dtype=np.float64, seed=1234),   # IRRG + nDSM (isprs:194-230)
    "potsdam": dict(H=6000, W=6000, C=5, K=6, dtype=np.float64, seed=1235),     # RGB + IR + nDSM
    "contest": dict(H=3989, W=2830, C=3, K=7, dtype=np.float32, seed=1236),     # GRSS-DFC2014 visible (contest:148-150)
    "coffee": dict(H=500, W=500, C=3, K=2, dtype=np.float32, seed=1237),        # coffee tiles (coffee:85-86)
}


def scene(kind, H=None, W=None, seed=None, block=50, unlabelled=None):
    """(image [H,W,C] in the dataset's dtype, labels [H,W] uint8)."""
    sp = SHAPES[kind]
    H = sp["H"] if H is None else H
    W = sp["W"] if W is None else W
    rs = np.random.RandomState(sp["seed"] if seed is None else seed)
    img = rs.randint(0, 256, size=(H, W, sp["C"]), dtype=np.uint8)
    img = (img / 255.0).astype(sp["dtype"]) if sp["dtype"] == np.float64 else (img.astype(np.float32) / np.float32(255.0))
    nb_h, nb_w = (H + block - 1) // block, (W + block - 1) // block
    k_draw = sp["K"] + (1 if unlabelled else 0)
    cls = rs.randint(0, k_draw, size=(nb_h, nb_w))
    lab = np.repeat(np.repeat(cls, block, axis=0), block, axis=1)[:H, :W].astype(np.uint8)
    return img, lab


def normalisation(img):
    """Per-channel mean / std stand-ins for dynamically_calculate_mean_and_std (isprs:151-184); channels 0..2 used."""
    flat = img.reshape(-1, img.shape[-1])[::97].astype(np.float64)
    return flat.mean(0), flat.std(0, ddof=1)


def random_instances(rs, n, shapes, crop_ref=25):
    """(map, row, col, rot) rows like select_super_batch_instances (isprs:403-445) without the class balancing."""
    out = np.zeros((n, 4), dtype=np.int64)
    out[:, 0] = rs.randint(0, len(shapes), size=n)
    for i in range(n):
        h, w = shapes[out[i, 0]]
        out[i, 1] = rs.randint(0, h - crop_ref + 1)
        out[i, 2] = rs.randint(0, w - crop_ref + 1)
    out[:, 3] = rs.randint(0, 360, size=n)
    return out
