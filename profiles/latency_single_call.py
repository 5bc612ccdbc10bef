"""Latency of the drop-in call depth_to_point_cloud(image, depth, ...) with NumPy inputs/outputs."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from tests import cases  # noqa: E402

for name, (H, W, h, w, dens) in {
    "480p dav2 medium (UI default)": (480, 640, 518, 686, "medium"),
    "480p dav2 high": (480, 640, 518, 686, "high"),
    "1080p dav2 medium": (1080, 1920, 518, 924, "medium"),
    "1080p dav2 high": (1080, 1920, 518, 924, "high"),
    "1080p native high": (1080, 1920, 1080, 1920, "high"),
    "4K dav2 high": (2160, 3840, 518, 924, "high"),
}.items():
    img = cases.make_image(H, W, 1)
    dep = cases.make_depth(h, w, 1, "uniform")
    for _ in range(3):
        p, c = m.depth_to_point_cloud(img, dep, density=dens)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        p, c = m.depth_to_point_cloud(img, dep, density=dens)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"{name:32s} points {len(p):8d}  median {ts[5]*1e3:7.2f} ms  min {ts[0]*1e3:7.2f} ms  -> {len(p)/ts[5]/1e6:8.1f} Mpoints/s")
