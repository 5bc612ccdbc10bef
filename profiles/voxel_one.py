"""One voxel-stage call on a 4K frame (for ncu): python profiles/voxel_one.py uniform|scene VOXEL_SIZE [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from image_to_pointcloud_b200.engine import EmitResult  # noqa: E402
from profiles.voxel_sweep import depth_maps  # noqa: E402

name, vs = sys.argv[1], float(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
H, W = 2160, 3840
dev = torch.device("cuda", 0)
maps, g = depth_maps(dev, H, W)
bgr = torch.randint(0, 256, (1, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
eng = m.FrameEngine(H, W, batch=1, device=dev)
cfg = eng.make_config(density="high", z_range=(0.5, 9.5), want_bounds=True)
xyz, rgb = eng.alloc_outputs(cfg)
cnt = torch.zeros(1, dtype=torch.int32, device=dev)
bounds = torch.empty((1, 6), dtype=torch.float32, device=dev)
s = torch.cuda.current_stream(dev)
eng.enqueue_stats(cfg, maps[name], s)
eng.enqueue_emit(cfg, maps[name], bgr, xyz, rgb, cnt, bounds, s)
res = EmitResult(xyz, rgb, cnt, bounds)
for _ in range(iters):
    vx, vr, vi, vc = eng.voxel_downsample(cfg, res, vs, check_error=False)
torch.cuda.synchronize()
print(name, vs, "points", int(cnt[0]), "voxels", int(vc[0]))
if os.environ.get("TIMELINE"):
    import json
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.voxel_downsample(cfg, res, vs, check_error=False)
        torch.cuda.synchronize()
    prof.export_chrome_trace("/tmp/_vox.json")
    ev = sorted((e for e in json.load(open("/tmp/_vox.json"))["traceEvents"] if e.get("cat") == "kernel"), key=lambda e: e["ts"])
    for e in ev:
        print("  %9.1f %8.1f  %-40s %s" % (e["ts"] - ev[0]["ts"], e["dur"], e["name"].split("(")[0].replace("d2pc::", "")[:40], e["args"].get("grid")))
