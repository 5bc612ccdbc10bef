"""One SOR call on a 1080p cloud (for ncu): python profiles/sor_one.py scene|uniform"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from profiles.voxel_sweep import depth_maps  # noqa: E402

kind = sys.argv[1]
H, W = 1080, 1920
dev = torch.device("cuda", 0)
maps, g = depth_maps(dev, H, W)
bgr = torch.randint(0, 256, (1, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
eng = m.FrameEngine(H, W, batch=1, device=dev)
res = eng.process(eng.make_config(density="high"), maps[kind], bgr)
p, c, idx, st = m.statistical_outlier_removal(res.xyz[0], res.rgb[0], return_device=True)
torch.cuda.synchronize()
print(kind, int(p.shape[0]), st)
if os.environ.get("TIMELINE"):  # per-kernel durations of a few calls (CUPTI through torch.profiler)
    import json
    from torch.profiler import ProfilerActivity, profile
    for it in range(int(os.environ["TIMELINE"])):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            m.statistical_outlier_removal(res.xyz[0], res.rgb[0], return_device=True)
            torch.cuda.synchronize()
        prof.export_chrome_trace("/tmp/_sor_tl.json")
        ev = sorted((e for e in json.load(open("/tmp/_sor_tl.json"))["traceEvents"] if e.get("cat") == "kernel"),
                    key=lambda e: e["ts"])
        tot = {}
        for e in ev:
            nm = e["name"].split("(")[0].replace("void d2pc::", "").replace("d2pc::", "")[:40]
            tot[nm] = tot.get(nm, 0.0) + e["dur"]
        print("call %d: span %.1f us; " % (it, ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]) +
              ", ".join("%s %.0f" % kv for kv in sorted(tot.items(), key=lambda kv: -kv[1])[:6]))
