"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/ from the UNMODIFIED reference function.

Run in the build container (needs /root/reference):  ``python -m oracle.make_golden``

* ``tests/golden/small_cases.npz``: inputs and full outputs of the reference's
  ``depth_to_point_cloud`` (``backend/app.py:174-250``) on small frames covering every branch
  (percentile / min-max fallback / zeros, non-finite repair, invert on/off, strides 1/2/4 on sizes
  that are not multiples of the stride, up- and down-scaling resizes, 2x2 sources, fov, 4-channel
  and grey images).
* ``tests/golden/large_cases.json``: for the BASELINE.json configurations (640x480, 1080p, 4K)
  the SHA-256 of the reference's output arrays, a strided sample of rows and the percentile
  scalars; inputs are regenerated from seeds by ``tests/cases.py``.

Every file records numpy / cv2 versions and ``cv2.ipp.useIPP()`` because the reference's
numerics depend on them (float64 chain under NumPy >= 2, IPP bilinear).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import reference_depth_to_point_cloud  # noqa: E402
from tests import cases  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def versions():
    import cv2
    return {"numpy": np.__version__, "cv2": cv2.__version__, "ipp": bool(cv2.ipp.useIPP()),
            "ipp_version": cv2.ipp.getIppVersion()}


def small_cases():
    """name -> (image, depth, kwargs)"""
    rng = np.random.default_rng(2024)
    out = {}
    img = rng.integers(0, 256, (24, 32, 3), dtype=np.uint8)
    dep_lo = (rng.random((19, 27)) * 20).astype(np.float32)
    dep_hi = (rng.random((61, 47)) * 20).astype(np.float32)
    dep_nat = (rng.random((24, 32)) * 20).astype(np.float32)
    for dens in ("low", "medium", "high"):
        for inv in (True, False):
            out[f"up_{dens}_inv{int(inv)}"] = (img, dep_lo, dict(density=dens, invert=inv))
    out["down_high"] = (img, dep_hi, dict(density="high"))
    out["native_high"] = (img, dep_nat, dict(density="high"))
    out["native_medium_scale15_noinv"] = (img, dep_nat, dict(density="medium", invert=False, depth_scale=15.0))
    out["fov60"] = (img, dep_lo, dict(density="high", fov=60.0))
    img_odd = rng.integers(0, 256, (31, 45, 3), dtype=np.uint8)
    dep_odd = (rng.random((23, 29)) * 5).astype(np.float32)
    for dens in ("low", "medium", "high"):
        out[f"odd_{dens}"] = (img_odd, dep_odd, dict(density=dens))
    d = dep_nat.copy(); d[3, 4] = np.nan; d[10, 10] = np.inf; d[20, 7] = -np.inf
    out["nonfinite_native_odd_count"] = (img, d, dict(density="high"))
    d = dep_nat.copy(); d[3, 4] = np.nan; d[10, 10] = np.inf
    out["nonfinite_native_even_count"] = (img, d, dict(density="high", invert=False))
    d = dep_lo.copy(); d[3, 4] = np.nan; d[10, 10] = np.inf; d[15, 7] = -np.inf
    out["nonfinite_resized"] = (img, d, dict(density="high"))
    out["constant_zeros"] = (img, np.full((24, 32), 3.0, np.float32), dict(density="high"))
    out["constant_zeros_noinv"] = (img, np.full((24, 32), 3.0, np.float32), dict(density="high", invert=False))
    d = np.full((24, 32), 3.0, np.float32); d[0, 0] = 1.0; d[5, 5] = 7.0
    out["minmax_fallback"] = (img, d, dict(density="high"))
    out["minmax_fallback_noinv"] = (img, d.copy(), dict(density="medium", invert=False))
    out["all_nan"] = (img, np.full((24, 32), np.nan, np.float32), dict(density="high"))
    d = dep_nat.copy(); d[:15] = np.inf
    out["majority_inf"] = (img, d, dict(density="high"))
    d = np.round(dep_nat).astype(np.float32); d[:6] = 0.0
    out["ties"] = (img, d, dict(density="high"))
    out["ties_noinv"] = (img, d.copy(), dict(density="high", invert=False))
    out["src_2x2"] = (img, (rng.random((2, 2)) * 20).astype(np.float32), dict(density="high"))
    out["bgra"] = (rng.integers(0, 256, (24, 32, 4), dtype=np.uint8), dep_lo, dict(density="high"))
    out["grey"] = (rng.integers(0, 256, (24, 32), dtype=np.uint8), dep_lo, dict(density="high"))
    out["neg_values"] = (img, (rng.standard_normal((24, 32)) * 100).astype(np.float32), dict(density="high"))
    out["tiny_1x1_img"] = (rng.integers(0, 256, (1, 1, 3), dtype=np.uint8),
                          np.array([[2.5]], np.float32), dict(density="high"))
    # a6: smooth=True (cv2.GaussianBlur on the normalised map), appended last so that the random
    # stream of the cases above is unchanged
    out["smooth_k5_up"] = (img, dep_lo, dict(density="high", smooth=True))
    out["smooth_k5_native_noinv"] = (img, dep_nat, dict(density="medium", smooth=True, invert=False))
    out["smooth_k3"] = (img, dep_nat, dict(density="high", smooth=True, smooth_ksize=3))
    out["smooth_k7_odd"] = (img_odd, dep_odd, dict(density="high", smooth=True, smooth_ksize=7))
    out["smooth_k9_odd_low"] = (img_odd, dep_odd, dict(density="low", smooth=True, smooth_ksize=8))
    out["smooth_zeros"] = (img, np.full((24, 32), 3.0, np.float32), dict(density="high", smooth=True))
    d = np.full((24, 32), 3.0, np.float32); d[0, 0] = 1.0; d[5, 5] = 7.0
    out["smooth_minmax_f32"] = (img, d, dict(density="high", smooth=True))
    # depth maps with a 1-pixel side: cv2.resize leaves IPP for OpenCV's own two-pass code (own random stream)
    r1 = np.random.default_rng(4242)
    out["src_1xN_up"] = (img, (r1.random((1, 11)) * 20).astype(np.float32), dict(density="high"))
    out["src_Nx1_noinv"] = (img, (r1.random((37, 1)) * 20).astype(np.float32), dict(density="medium", invert=False))
    out["src_1x1_resized"] = (img, np.array([[2.5]], np.float32), dict(density="high"))
    d = (r1.standard_normal((1, 45)) * 5).astype(np.float32); d[0, 7] = np.inf; d[0, 30] = np.nan
    out["src_1xN_nonfinite_down"] = (img_odd, d, dict(density="high"))
    d = (r1.random((40, 1)) * 9).astype(np.float32); d[3, 0] = -np.inf
    out["src_Nx1_nonfinite_low"] = (img_odd, d, dict(density="low"))
    return out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = reference_depth_to_point_cloud()
    ver = versions()
    assert ver["numpy"].startswith("2."), ver
    store = {"__versions__": np.array(json.dumps(ver))}
    names = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, (img, dep, kw) in small_cases().items():
            pts, cols = ref(img, dep, **kw)
            store[f"{name}/image"] = img
            store[f"{name}/depth"] = dep
            store[f"{name}/kwargs"] = np.array(json.dumps(kw))
            store[f"{name}/points"] = pts
            store[f"{name}/colors"] = cols
            names.append(name)
            print(f"small {name:32s} N={len(pts)}")
    # f2: depth preview data URLs of the reference's create_depth_preview (app.py:124-172)
    from oracle.ref_loader import load_reference_module
    refmod = load_reference_module()
    prng = np.random.default_rng(77)
    pnames = []
    for pname, pd_, inv in [("preview_uniform", (prng.random((40, 56)) * 20).astype(np.float32), True),
                            ("preview_noinv", (prng.standard_normal((33, 47)) * 3).astype(np.float32), False),
                            ("preview_constant", np.full((16, 24), 2.0, np.float32), True)]:
        store[f"{pname}/depth"] = pd_
        store[f"{pname}/invert"] = np.array(inv)
        store[f"{pname}/data_url"] = np.array(refmod.create_depth_preview(pd_, invert=inv))
        pnames.append(pname)
        print(f"preview {pname}")
    store["__preview_names__"] = np.array(json.dumps(pnames))
    store["__names__"] = np.array(json.dumps(names))
    np.savez_compressed(os.path.join(GOLDEN, "small_cases.npz"), **store)

    large = {"versions": ver, "cases": {}}
    only = set(sys.argv[1:])
    path = os.path.join(GOLDEN, "large_cases.json")
    if only and os.path.exists(path):
        large = json.load(open(path))
    for name, spec in cases.LARGE_CASES.items():
        if only and name not in only:
            continue
        img, dep, kw = cases.build_case(spec)
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pts, cols = ref(img, dep, **kw)
        dt = time.perf_counter() - t0
        n = len(pts)
        stride = max(1, n // 64)
        large["cases"][name] = {
            "spec": {"img": list(spec["img"]), "depth": list(spec["depth"]), "kw": kw},
            "n_points": n,
            "image_sha256": sha(img), "depth_sha256": sha(dep),
            "points_sha256": sha(pts), "colors_sha256": sha(cols),
            "sample_stride": stride,
            "points_sample_hex": pts[::stride].tobytes().hex(),
            "colors_sample_hex": cols[::stride].tobytes().hex(),
            "reference_seconds": round(dt, 3),
        }
        print(f"large {name:28s} N={n} ref {dt:.2f}s")
        json.dump(large, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
