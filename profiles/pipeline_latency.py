"""app.py:468-559 (stage -> SOR -> preview -> PLY bytes -> bounds): one device-resident call vs the same
steps as separate drop-ins chained through host arrays.  NumPy image + depth in, everything the job result
needs out.   python profiles/pipeline_latency.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from tests import cases  # noqa: E402


def timeit(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return sorted(ts)[n // 2] * 1e3


out = {}
os.makedirs("/tmp/d2pc_out/outputs", exist_ok=True)
os.chdir("/tmp/d2pc_out")
for name, (H, W, h, w, dens) in {"480p_medium (UI default)": (480, 640, 518, 686, "medium"),
                                 "1080p_medium": (1080, 1920, 518, 924, "medium"),
                                 "1080p_high": (1080, 1920, 518, 924, "high")}.items():
    img = cases.make_image(H, W, 1)
    dep = cases.make_depth(h, w, 1, "scene")

    def fused():
        return m.point_cloud_stage(img, dep, density=dens, output_format="ply", filename="a")

    def chained():
        p, c = m.depth_to_point_cloud(img, dep, density=dens)
        p, c = m.refine_point_cloud(p, c)
        pp, pc = m.preview_lists(p, c)
        path = m.save_point_cloud(p, c, "ply", "b")
        return p[:, 0].min(), path

    r = fused()
    out[name] = {"points_kept": r["point_count"], "fused_ms": round(timeit(fused), 2), "chained_drop_ins_ms": round(timeit(chained), 2)}
print(json.dumps({"workload": "stage + SOR + preview + PLY file + bounds, NumPy in, file on disk out", "pipeline": out}))
