"""Randomised check of row f3 (preview rows, XYZ text, LAS and PLY records) on the device against the oracle:
python profiles/fuzz_writers.py [runs] [seed].  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from oracle import d2pc_oracle as O  # noqa: E402  (checker only)


def rows(rng, n, kind):
    if kind == "bits":       # random float32 bit patterns, finite, |x| < 2^43 (what "%.6f" prints exactly)
        v = rng.integers(0, 1 << 32, (n * 4, 3), dtype=np.uint64).astype(np.uint32).view(np.float32)
        ok = np.isfinite(v).all(axis=1) & (np.abs(v) < 2.0 ** 43).all(axis=1)
        v = v[ok][:n]
    elif kind == "ties":     # decimal ties of "%.6f": odd multiples of 2^-k
        v = ((rng.integers(0, 1 << 20, (n, 3)) * 2 + 1) / 2.0 ** rng.integers(1, 30, (n, 3))).astype(np.float32)
        v *= rng.choice([-1.0, 1.0], (n, 3)).astype(np.float32)
    elif kind == "cloud":    # what the stage emits
        v = (rng.standard_normal((n, 3)) * rng.choice([0.01, 1.0, 10.0])).astype(np.float32)
    else:                    # tiny: denormals, signed zeros
        v = (rng.standard_normal((n, 3)) * 1e-38).astype(np.float32)
        v[rng.random((n, 3)) < 0.1] = -0.0
    c = rng.integers(0, 256, (len(v), 3)).astype(np.float32)
    return np.ascontiguousarray(v, dtype=np.float32), c


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    runs = int(argv[1]) if len(argv) > 1 else 40
    seed = int(argv[2]) if len(argv) > 2 else 5
    rng = np.random.default_rng(seed)
    t0 = time.time()
    res = {"xyz_text": [0, 0], "las": [0, 0], "ply": [0, 0], "preview": [0, 0]}   # [runs, mismatches]
    n_rows = 0
    for it in range(runs):
        kind = ["bits", "ties", "cloud", "tiny"][it % 4]
        n = int(rng.choice([1, 255, 257, 4097, 20001, 60000, 150000]))
        p, c = rows(rng, n, kind)
        n_rows += len(p)
        res["xyz_text"][0] += 1
        res["xyz_text"][1] += m.xyz_text(p, c) != O.xyz_text(p, c)
        gp, gc = m.preview_rows(p, c)
        wp, wc = O.preview_rows(p, c)
        res["preview"][0] += 1
        res["preview"][1] += not (np.array_equal(gp.view(np.uint32), wp.view(np.uint32)) and np.array_equal(gc, wc))
        res["ply"][0] += 1
        res["ply"][1] += m.ply_vertex_records(p, c).tobytes() != O.ply_records(p, c).tobytes()
        if kind in ("cloud", "tiny", "ties") and len(p) and float(np.abs(p).max()) < 1e6:
            rec, off, mm = m.las_point_records(p, c)
            want, woff = O.las_records(p, c)
            res["las"][0] += 1
            res["las"][1] += not (off == woff and rec.tobytes() == want.tobytes())
    out = {k: {"runs": v[0], "mismatches": int(v[1])} for k, v in res.items()}
    print(json.dumps({"seed": seed, "rows": int(n_rows), **out, "seconds": round(time.time() - t0, 1)}))
    return 1 if any(v[1] for v in res.values()) else 0


if __name__ == "__main__":
    sys.exit(main())
