"""BASELINE.json configs[2] and [4]: 4K frame, depth-range mask 0.5..9.5, voxel-size sweep 1 mm - 5 cm.
Times the voxel stage alone (CUDA events, device-resident emitted rows) for the two synthetic depth
distributions of SURVEY.md 8d: (i) uniform*20 (every point its own voxel at small sizes: worst case),
(ii) smooth scene 20/(1+r) + noise (voxels merge).  Prints one JSON object.

    python profiles/voxel_sweep.py [iters]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from image_to_pointcloud_b200.engine import EmitResult  # noqa: E402

SIZES = [0.001, 0.002, 0.005, 0.01, 0.02, 0.05]


def depth_maps(dev, H, W):
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    uni = torch.rand((1, H, W), generator=g, device=dev) * 20
    vv, uu = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    r = torch.hypot(uu - W / 2.0, vv - H / 2.0) / W
    scene = (20.0 / (1.0 + r) + torch.randn((H, W), generator=g, device=dev) * 0.01).float().reshape(1, H, W).contiguous()
    return {"uniform": uni, "scene": scene}, g


def sweep(H=2160, W=3840, iters=5, peak=None):
    dev = torch.device("cuda", 0)
    maps, g = depth_maps(dev, H, W)
    bgr = torch.randint(0, 256, (1, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
    eng = m.FrameEngine(H, W, batch=1, device=dev)
    cfg = eng.make_config(density="high", z_range=(0.5, 9.5), want_bounds=True)
    xyz, rgb = eng.alloc_outputs(cfg)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    bounds = torch.empty((1, 6), dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for name, depth in maps.items():
        eng.enqueue_stats(cfg, depth, s)
        eng.enqueue_emit(cfg, depth, bgr, xyz, rgb, cnt, bounds, s)
        res = EmitResult(xyz, rgb, cnt, bounds)
        M = int(cnt[0])
        rows = {}
        for vs in SIZES:
            for _ in range(2):
                vx, vr, vi, vc = eng.voxel_downsample(cfg, res, vs, check_error=False)
            torch.cuda.synchronize()
            a.record()
            for _ in range(iters):
                vx, vr, vi, vc = eng.voxel_downsample(cfg, res, vs, check_error=False)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            V = int(vc[0])
            alg = 24 * M + 24 * V
            rows[str(vs)] = {"ms": round(ms, 4), "voxels": V, "mpoints_in_per_s": round(M / ms / 1e3, 1),
                             "alg_gbs": round(alg / ms / 1e6, 1)}
            if peak:
                rows[str(vs)]["frac_of_peak"] = round(alg / ms / 1e6 / peak, 4)
        out[name] = {"points_in": M, "sizes": rows}
    return out


def config2(H=2160, W=3840, voxel=0.005, iters=5, peak=None):
    """BASELINE configs[2]: ONE 4K frame (smooth scene), statistics + z-range-masked emit + 5 mm voxel grid,
    device-resident, CUDA-event timed; algorithmic bytes 4 D + 3 N + 24 M (emit) + 24 M + 24 V (voxel stage)."""
    dev = torch.device("cuda", 0)
    maps, g = depth_maps(dev, H, W)
    depth = maps["scene"]
    bgr = torch.randint(0, 256, (1, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
    eng = m.FrameEngine(H, W, batch=1, device=dev)
    cfg = eng.make_config(density="high", z_range=(0.5, 9.5), want_bounds=True)
    xyz, rgb = eng.alloc_outputs(cfg)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    bounds = torch.empty((1, 6), dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream(dev)
    res = EmitResult(xyz, rgb, cnt, bounds)

    def stage():
        eng.enqueue_path(cfg, depth, bgr, xyz, rgb, cnt, bounds, s)

    def whole():
        stage()
        return eng.voxel_downsample(cfg, res, voxel, check_error=False)
    out = {}
    for name, fn in (("stage_ms", stage), ("whole_ms", whole)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            r = fn()
        b.record()
        torch.cuda.synchronize()
        out[name] = round(a.elapsed_time(b) / iters, 4)
    M, V = int(cnt[0]), int(r[3][0])
    alg = 4 * H * W + 3 * H * W + 24 * M + 24 * M + 24 * V
    out.update({"points_in": H * W, "kept": M, "voxels": V, "voxel_size": voxel, "alg_bytes": alg,
                "alg_gbs": round(alg / out["whole_ms"] / 1e6, 1)})
    if peak:
        out["frac_of_measured_peak"] = round(alg / out["whole_ms"] / 1e6 / peak, 4)
    return out


if __name__ == "__main__":
    it = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    peak = None
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    print(json.dumps({"workload": "4K frame, z_range 0.5..9.5, voxel stage only (insert + extract)", "sweep": sweep(iters=it, peak=peak)}))
