"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py
(C1..C5).  Everything is a pure function of the seed
(``np.random.default_rng`` = PCG64, whose stream NumPy keeps stable across versions)."""
from __future__ import annotations

import numpy as np


def make_image(h: int, w: int, seed: int) -> np.ndarray:
    """BGR uint8 [h, w, 3], like ``cv2.imdecode(..., IMREAD_COLOR)`` (reference app.py:433)."""
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def make_depth(h: int, w: int, seed: int, kind: str = "uniform") -> np.ndarray:
    """float32 [h, w] raw 'network output' (larger = nearer)."""
    rng = np.random.default_rng(seed + 7919)
    if kind == "uniform":
        return (rng.random((h, w)) * 20).astype(np.float32)
    if kind == "scene":
        # smooth scene so that neighbouring pixels share voxels: 20/(1+|uv-c|/w) + noise
        v, u = np.mgrid[0:h, 0:w].astype(np.float64)
        r = np.hypot(u - w / 2.0, v - h / 2.0) / w
        d = 20.0 / (1.0 + r) + rng.standard_normal((h, w)) * 0.01
        return d.astype(np.float32)
    if kind == "nonfinite":
        d = (rng.random((h, w)) * 20).astype(np.float32)
        n = h * w
        idx = rng.choice(n, max(3, n // 100), replace=False)
        vals = np.array([np.nan, np.inf, -np.inf], dtype=np.float32)
        d.ravel()[idx] = vals[np.arange(idx.size) % 3]
        return d
    if kind == "ties":
        # heavy ties: quantised values + a saturated 'sky' region
        d = np.round(rng.random((h, w)) * 20).astype(np.float32)
        d[: h // 5] = 0.0
        return d
    if kind == "constant":
        return np.full((h, w), 3.0, dtype=np.float32)
    if kind == "two_outliers":
        d = np.full((h, w), 3.0, dtype=np.float32)
        d[0, 0] = 1.0
        d[h // 2, w // 2] = 7.0
        return d
    raise KeyError(kind)


# name -> dict(img=(H,W,seed), depth=(h,w,seed,kind), kwargs)
LARGE_CASES = {
    "c1_480p_high": dict(img=(480, 640, 0), depth=(518, 686, 0, "uniform"), kw=dict(density="high")),
    "c1_480p_medium": dict(img=(480, 640, 0), depth=(518, 686, 0, "uniform"), kw=dict(density="medium")),
    "c1_480p_low_noinv": dict(img=(480, 640, 0), depth=(518, 686, 0, "uniform"),
                              kw=dict(density="low", invert=False, depth_scale=15.0)),
    "c2_1080p_native": dict(img=(1080, 1920, 1), depth=(1080, 1920, 1, "uniform"), kw=dict(density="high")),
    "c2_1080p_dav2": dict(img=(1080, 1920, 1), depth=(518, 924, 1, "uniform"), kw=dict(density="high")),
    "c2_1080p_scene_nonfinite": dict(img=(1080, 1920, 1), depth=(518, 924, 11, "nonfinite"),
                                     kw=dict(density="medium")),
    "c3_4k_dav2": dict(img=(2160, 3840, 2), depth=(518, 924, 2, "scene"), kw=dict(density="high")),
}


def build_case(spec):
    H, W, s = spec["img"]
    h, w, ds, kind = spec["depth"]
    return make_image(H, W, s), make_depth(h, w, ds, kind), dict(spec["kw"])


def writer_rows(seed: int = 7, n: int = 4000):
    """Rows that stress the writers' formatting (row f3): decimal ties (odd j / 128 has a 7th decimal of
    exactly 5), values that round to +-0.000000, signed zeros, large and tiny magnitudes, denormals."""
    rng = np.random.default_rng(seed)
    p = (rng.standard_normal((n, 3)) * 10).astype(np.float32)
    special = np.array([0.0, -0.0, 1e-9, -1e-9, 4.9e-7, 5.1e-7, -5e-7, 1 / 128, 3 / 128, -5 / 128, 7 / 256, 0.5, -0.5,
                        1.0, -1.0, 9.9999995, 123456.789, -98765.4321, 1e7, 16777216.0, 3.0e12, 2.0 ** 43,
                        1e-38, 1e-45, -1e-45, 0.1, 0.2, 0.3, 2.675, 1.0000005, 0.9999995, 999999.5],
                       dtype=np.float32)
    k = len(special)
    p[:k, 0] = special
    p[:k, 1] = special[::-1]
    p[:k, 2] = np.roll(special, 5)
    p[k:2 * k, :] = (np.arange(k * 3, dtype=np.float32).reshape(k, 3) * 2 + 1) / 128   # more exact ties
    c = rng.integers(0, 256, (n, 3)).astype(np.float32)
    c[:4] = [[0, 0, 0], [255, 255, 255], [128, 128, 128], [1, 254, 17]]
    return p, c
