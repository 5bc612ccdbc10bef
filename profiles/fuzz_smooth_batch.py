"""Randomised check of a6 (smooth=True) and of the batch API against the oracle:
python profiles/fuzz_smooth_batch.py [examples] [seed].  Prints one JSON line."""
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


def oracle(img, dep, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return O.depth_to_point_cloud(img, dep, **kw)


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    n = int(argv[1]) if len(argv) > 1 else 300
    seed = int(argv[2]) if len(argv) > 2 else 8
    rng = np.random.default_rng(seed)
    t0 = time.time()
    sm = {"examples": 0, "bit_exact": 0, "within_1e-5": 0, "mismatches": 0}
    bt = {"batches": 0, "frames": 0, "mismatches": 0}
    bad = []
    for it in range(n):
        H, W = int(rng.integers(1, 200)), int(rng.integers(1, 280))
        native = rng.random() < 0.4
        h, w = (H, W) if native else (int(rng.integers(1, 160)), int(rng.integers(1, 220)))
        if not native and rng.random() < 0.1:   # depth maps with a 1-pixel side (OpenCV's non-IPP resize)
            h, w = (1, w) if rng.random() < 0.5 else (h, 1)
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        dist = DISTS[int(rng.integers(len(DISTS)))]
        dep = _depth(rng, h, w, dist)
        kw = dict(density=str(rng.choice(["low", "medium", "high"])), invert=bool(rng.integers(2)),
                  depth_scale=float(rng.choice([10.0, 1.0, 250.0])), smooth=True,
                  smooth_ksize=int(rng.choice([3, 4, 5, 5, 6, 7, 8, 9])))
        po, co = oracle(img, dep, **kw)
        p, c = m.depth_to_point_cloud(img, dep, **kw)
        sm["examples"] += 1
        if p.tobytes() == po.tobytes() and c.tobytes() == co.tobytes():
            sm["bit_exact"] += 1
        elif p.shape == po.shape and np.allclose(p, po, rtol=1e-5, atol=1e-6, equal_nan=True) and c.tobytes() == co.tobytes():
            sm["within_1e-5"] += 1
        else:
            sm["mismatches"] += 1
            bad.append(dict(row="a6", it=it, H=H, W=W, h=h, w=w, dist=dist, kw=kw))
        if it % 4 == 0:   # a batch of frames of one geometry with mixed value distributions (fast and fallback frames)
            nb = int(rng.integers(1, 12))
            imgs = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(nb)]
            deps = [_depth(rng, h, w, DISTS[int(rng.integers(len(DISTS)))]) for _ in range(nb)]
            kwb = dict(density=kw["density"], invert=kw["invert"], depth_scale=kw["depth_scale"])
            zr = (1.0, 8.0) if rng.random() < 0.4 else None
            outs = m.depth_to_point_cloud_batch(imgs, deps, z_range=zr, chunk=int(rng.choice([1, 3, 8])), **kwb)
            bt["batches"] += 1
            for i in range(nb):
                po, co = oracle(imgs[i], deps[i], **kwb)
                if zr is not None:
                    keep = O.range_mask(po, *zr)
                    po, co = po[keep], co[keep]
                ok = outs[i][0].tobytes() == po.tobytes() and outs[i][1].tobytes() == co.tobytes()
                bt["frames"] += 1
                if not ok:
                    bt["mismatches"] += 1
                    bad.append(dict(row="batch", it=it, frame=i, H=H, W=W, h=h, w=w, kw=kwb, z_range=zr))
    print(json.dumps({"seed": seed, "a6_smooth": sm, "batch_api": bt, "seconds": round(time.time() - t0, 1), "first_bad": bad[:6]}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
