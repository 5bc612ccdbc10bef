"""Randomised parity run against the oracle (bit-exact): python profiles/fuzz_parity.py [examples] [seed] [big]
Shapes up to 300 x 420 (depth maps down to 1-pixel sides), every density / invert / depth_scale / fov, seven value
distributions, 1/3/4-channel images, optional depth-range mask and drop_nonfinite.  Prints one JSON line."""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from oracle import d2pc_oracle as O  # noqa: E402  (checker only)
from tests.test_property_gpu import DISTS, _depth  # noqa: E402


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    n = int(argv[1]) if len(argv) > 1 else 1000
    seed = int(argv[2]) if len(argv) > 2 else 2026
    big = len(argv) > 3 and argv[3] == "big"   # frames of 0.2 - 2.4 Mpixel: sampled statistics, many tiles
    rng = np.random.default_rng(seed)
    bad, points, t0 = [], 0, time.time()
    stats = {"resized": 0, "one_pixel_side": 0, "masked": 0, "nonfinite": 0, "empty_result": 0}
    for it in range(n):
        H, W = (int(rng.integers(400, 1200)), int(rng.integers(500, 2000))) if big else \
            (int(rng.integers(1, 300)), int(rng.integers(1, 420)))
        if rng.random() < 0.35:
            h, w = H, W
        else:
            h, w = (int(rng.integers(200, 1300)), int(rng.integers(200, 2100))) if big else \
                (int(rng.integers(1, 260)), int(rng.integers(1, 360)))
            if rng.random() < 0.1:
                h = 1
            elif rng.random() < 0.1:
                w = 1
        chans = int(rng.choice([3, 3, 3, 4, 1]))
        img = rng.integers(0, 256, (H, W) if chans == 1 else (H, W, chans), dtype=np.uint8)
        dist = DISTS[int(rng.integers(len(DISTS)))]
        dep = _depth(rng, h, w, dist)
        kw = dict(density=str(rng.choice(["low", "medium", "high"])), invert=bool(rng.integers(2)),
                  depth_scale=float(rng.choice([10.0, 1.0, 3.7, 250.0, -2.0])))
        if rng.random() < 0.2:
            kw["fov"] = float(rng.choice([30.0, 60.0, 90.0, 120.0]))
        zr = None
        if rng.random() < 0.4:
            a, b = sorted(rng.random(2) * abs(kw["depth_scale"]) * 1.1)
            zr = (float(a), float(b))
        dnf = bool(rng.random() < 0.2)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            po, co = O.depth_to_point_cloud(img, dep, **kw)
        keep = np.ones(len(po), bool)
        if zr is not None:
            keep &= O.range_mask(po, *zr)
        if dnf:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                dr = O.resize_bilinear(dep, H, W) if (h, w) != (H, W) else dep
            step = {"low": 4, "medium": 2, "high": 1}[kw["density"]]
            keep &= np.isfinite(dr[::step, ::step]).ravel()
        p, c = m.depth_to_point_cloud(img, dep, z_range=zr, drop_nonfinite=dnf, **kw)
        ok = (len(p) == int(keep.sum()) and p.tobytes() == po[keep].tobytes() and c.tobytes() == co[keep].tobytes())
        if not ok:
            bad.append(dict(it=it, H=H, W=W, h=h, w=w, chans=chans, dist=dist, kw=kw, z_range=zr, drop_nonfinite=dnf))
        points += len(p)
        stats["resized"] += (h, w) != (H, W)
        stats["one_pixel_side"] += ((h, w) != (H, W)) and (h == 1 or w == 1)
        stats["masked"] += zr is not None or dnf
        stats["nonfinite"] += dist == "nonfinite"
        stats["empty_result"] += len(p) == 0
    print(json.dumps({"examples": n, "seed": seed, "mismatches": len(bad), "points_compared": int(points),
                      "covered": {k: int(v) for k, v in stats.items()}, "seconds": round(time.time() - t0, 1),
                      "first_bad": bad[:5]}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
