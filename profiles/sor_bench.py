"""Row f1 timing: statistical outlier removal (k = 20, std_ratio = 2) on device-resident clouds produced by the
stage itself, CUDA events around d2pc_sor_enqueue (+ the bounds pass), and beside it the CPU baseline on the same
clouds: the oracle (scipy cKDTree k-NN + Open3D's statistics; Open3D itself is not installed) on 1 core and on all
host cores (cKDTree.query(workers=-1)).   python profiles/sor_bench.py [--no-cpu]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from profiles.voxel_sweep import depth_maps  # noqa: E402


def cpu_baseline(points_np, cores):
    """Seconds of the oracle's k-NN + statistics on `cores` host cores (-1 = all)."""
    import time

    import numpy as np
    from scipy.spatial import cKDTree
    p = np.ascontiguousarray(points_np, dtype=np.float64)
    t0 = time.perf_counter()
    dist, _ = cKDTree(p).query(p, k=20, workers=cores)
    avg = dist.sum(axis=1) / 20
    pos = avg > 0
    mean = avg[pos].sum() / len(p)
    std = np.sqrt((((avg - mean) ** 2) * pos).sum() / (len(p) - 1))
    keep = np.nonzero(pos & (avg < mean + 2.0 * std))[0]
    return time.perf_counter() - t0, len(keep)


def bench(iters=5, cpu=True):
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
            per_call = []
            for _ in range(iters):
                a.record()
                p, c, idx, st = m.statistical_outlier_removal(xyz, rgb, return_device=True)
                b.record()
                torch.cuda.synchronize()
                per_call.append(a.elapsed_time(b))
            ms = sorted(per_call)[len(per_call) // 2]   # median: a call is one synchronous request
            if os.environ.get("PER_CALL"):
                print(name, kind, ["%.2f" % x for x in per_call], file=sys.stderr)
            if os.environ.get("TIMELINE"):   # where a call's time goes (CUPTI through torch.profiler)
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    m.statistical_outlier_removal(xyz, rgb, return_device=True)
                    torch.cuda.synchronize()
                prof.export_chrome_trace("/tmp/_sorb.json")
                ev = sorted((e for e in json.load(open("/tmp/_sorb.json"))["traceEvents"] if e.get("cat") == "kernel"),
                            key=lambda e: e["ts"])
                tot = {}
                for e in ev:
                    nm = e["name"].split("(")[0].replace("void d2pc::", "").replace("d2pc::", "")[:32]
                    tot[nm] = tot.get(nm, 0.0) + e["dur"]
                print(name, kind, "%.2f ms timed; profiled call span %.0f us: " % (ms, ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]) +
                      ", ".join("%s %.0f" % kv for kv in sorted(tot.items(), key=lambda kv: -kv[1])[:5]), file=sys.stderr)
            row = {"points": int(xyz.shape[0]), "kept": int(p.shape[0]), "ms": round(ms, 3),
                   "mpoints_per_s": round(xyz.shape[0] / ms / 1e3, 1)}
            if cpu and name != "1080p_high" or (cpu and kind == "scene"):
                pts = xyz.cpu().numpy()
                t1, k1 = cpu_baseline(pts, 1)
                tn, kn = cpu_baseline(pts, -1)
                row["cpu_oracle_scipy"] = {"s_1_core": round(t1, 3), "s_all_cores": round(tn, 3), "cores": os.cpu_count(),
                                           "kept": k1, "speedup_vs_all_cores": round(tn * 1e3 / ms, 1)}
            out[f"{name}_{kind}"] = row
    return out


if __name__ == "__main__":
    print(json.dumps({"workload": "SOR k=20 std_ratio=2 on the stage's own clouds (scene / uniform-random depth)",
                      "sor": bench(cpu="--no-cpu" not in sys.argv)}))
