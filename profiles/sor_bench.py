"""Row f1 timing: statistical outlier removal (k = 20, std_ratio = 2) on device-resident clouds produced by the
stage itself, CUDA events around d2pc_sor_enqueue (+ the bounds pass).   python profiles/sor_bench.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from profiles.voxel_sweep import depth_maps  # noqa: E402


def bench(iters=5):
    dev = torch.device("cuda", 0)
    out = {}
    for name, (H, W, dens) in {"480p_medium": (480, 640, "medium"), "1080p_medium": (1080, 1920, "medium"),
                               "1080p_high": (1080, 1920, "high")}.items():
        maps, g = depth_maps(dev, H, W)
        bgr = torch.randint(0, 256, (1, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
        eng = m.FrameEngine(H, W, batch=1, device=dev)
        cfg = eng.make_config(density=dens)
        for kind, depth in maps.items():
            res = eng.process(cfg, depth, bgr)
            xyz, rgb = res.xyz[0], res.rgb[0]
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(2):   # allocator and first-launch warm-up
                m.statistical_outlier_removal(xyz, rgb, return_device=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(iters):
                p, c, idx, st = m.statistical_outlier_removal(xyz, rgb, return_device=True)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            out[f"{name}_{kind}"] = {"points": int(xyz.shape[0]), "kept": int(p.shape[0]), "ms": round(ms, 3),
                                     "mpoints_per_s": round(xyz.shape[0] / ms / 1e3, 1)}
    return out


if __name__ == "__main__":
    print(json.dumps({"workload": "SOR k=20 std_ratio=2 on the stage's own clouds (scene / uniform-random depth)", "sor": bench()}))
