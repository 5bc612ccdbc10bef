"""Runs one workload variant a few times (for ncu captures): python profiles/run_variant.py NAME [iters]
NAME: native | dav2 | mask | mask4k | native4k | medium_dav2 | scene_mask"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402

VARIANTS = {
    "native": (1080, 1920, 1080, 1920, 16, None, "high"),
    "dav2": (1080, 1920, 518, 924, 16, None, "high"),
    "mask": (1080, 1920, 1080, 1920, 16, (0.5, 9.5), "high"),
    "mask4k": (2160, 3840, 2160, 3840, 8, (0.5, 9.5), "high"),
    "native4k": (2160, 3840, 2160, 3840, 8, None, "high"),
    "medium_dav2": (1080, 1920, 518, 924, 16, None, "medium"),
    "mask4k_1": (2160, 3840, 2160, 3840, 1, (0.5, 9.5), "high"),
    "scene4k_1": (2160, 3840, 2160, 3840, 1, (0.5, 9.5), "high"),
    "native128": (1080, 1920, 1080, 1920, 128, None, "high"),
    "native4k16": (2160, 3840, 2160, 3840, 16, None, "high"),
    "mask4k16": (2160, 3840, 2160, 3840, 16, (0.5, 9.5), "high"),
    "dav2_64": (1080, 1920, 518, 924, 64, None, "high"),
    "mask64": (1080, 1920, 1080, 1920, 64, (0.5, 9.5), "high"),
    "mask_medium64": (1080, 1920, 1080, 1920, 64, (0.5, 9.5), "medium"),
    "medium_dav2_64": (1080, 1920, 518, 924, 64, None, "medium"),
    "scene_mask": (1080, 1920, 1080, 1920, 16, (2.0, 6.0), "high"),   # coherent mask: whole tiles kept / dropped
}


def main():
    name = sys.argv[1]
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    H, W, h, w, B, zr, dens = VARIANTS[name]
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    eng = m.FrameEngine(H, W, h, w, batch=B, device=dev)
    cfg = eng.make_config(density=dens, z_range=zr)
    depth = torch.rand((B, h, w), generator=g, device=dev) * 20
    if name.startswith("scene"):  # smooth scene so that voxels actually merge (SURVEY 8d distribution ii)
        vv, uu = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
        r = torch.hypot(uu - w / 2.0, vv - h / 2.0) / w
        depth = (20.0 / (1.0 + r) + torch.randn((h, w), generator=g, device=dev) * 0.01).float().expand(B, h, w).contiguous()
    bgr = torch.randint(0, 256, (B, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
    xyz, rgb = eng.alloc_outputs(cfg)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(iters + 2):
        if i == 2:
            a.record()
        eng.enqueue_stats(cfg, depth, s)
        eng.enqueue_emit(cfg, depth, bgr, xyz, rgb, cnt, None, s)
    b.record()
    torch.cuda.synchronize()
    print(name, "ms/iter", a.elapsed_time(b) / iters, "kept", int(cnt.sum()), "of", B * eng.points_per_frame(cfg))
    if os.environ.get("TIMELINE"):  # per-kernel start / duration of one iteration (CUPTI through torch.profiler)
        import json
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            eng.enqueue_stats(cfg, depth, s)
            eng.enqueue_emit(cfg, depth, bgr, xyz, rgb, cnt, None, s)
            torch.cuda.synchronize()
        path = "/tmp/_tl_%s.json" % name
        prof.export_chrome_trace(path)
        ev = sorted((e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"), key=lambda e: e["ts"])
        t0 = ev[0]["ts"]
        for e in ev:
            nm = e["name"].split("(")[0].replace("void d2pc::", "").replace("d2pc::", "")[:40]
            print("  %9.1f %8.1f  %-40s %s" % (e["ts"] - t0, e["dur"], nm, e["args"].get("grid")))
    if len(sys.argv) > 3:  # voxel sizes: time emit(with bounds)+voxel per size
        from image_to_pointcloud_b200.engine import EmitResult
        cfgb = eng.make_config(density=dens, z_range=zr, want_bounds=True)
        bounds = torch.empty((B, 6), dtype=torch.float32, device=dev)
        for vs in [float(x) for x in sys.argv[3].split(",")]:
            for i in range(3):
                if i == 1:
                    a.record()
                eng.enqueue_stats(cfgb, depth, s)
                eng.enqueue_emit(cfgb, depth, bgr, xyz, rgb, cnt, bounds, s)
                res = EmitResult(xyz, rgb, cnt, bounds)
                vx, vr, vi, vc = eng.voxel_downsample(cfgb, res, vs)
            b.record()
            torch.cuda.synchronize()
            print(f"  voxel {vs}: ms/iter {a.elapsed_time(b) / 2:.3f} (stats+emit+voxel, {B} frames) voxels {int(vc.sum())} of {int(cnt.sum())} points")


if __name__ == "__main__":
    main()
