"""Kernel timeline (CUPTI through torch.profiler) of one pipelined step: start / duration / stream of every
kernel, to see what overlaps and where the gaps are.  usage: python profiles/path_timeline.py S L [kind] [B]"""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402


def main(S=4, L=1, kind="native", B=32, graph=0, flags=0):
    dev = torch.device("cuda", 0)
    H, W = 1080, 1920
    h, w = (518, 924) if kind == "dav2" else (H, W)
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    depth = torch.rand((B, h, w), generator=g, device=dev) * 20
    bgr = torch.randint(0, 256, (B, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
    eng = m.FrameEngine(H, W, h, w, batch=B, device=dev)
    cfg = eng.make_config(density="high")
    xyz, rgb = eng.alloc_outputs(cfg)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream(dev)

    def step():
        eng.enqueue_path(cfg, depth, bgr, xyz, rgb, cnt, None, s, sub_batch=S, lookahead=L, graph=bool(graph), flags=flags)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    path = "gpurun_out/timeline_S%d_L%d_%s.json" % (S, L, kind)
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    rows = []
    for e in ev:
        name = e["name"].split("(")[0].replace("void d2pc::", "").replace("d2pc::", "")[:28]
        rows.append((e["ts"] - t0, e["dur"], e["args"].get("stream"), name, e["args"].get("grid")))
    with open(path.replace(".json", ".txt"), "w") as f:
        for r in rows:
            f.write("%9.1f %8.1f  s%-3s %-28s %s\n" % r)
    os.remove(path)
    end = max(r[0] + r[1] for r in rows)
    print("kernels %d, span %.1f us" % (len(rows), end))
    for r in rows[:60]:
        print("%9.1f %8.1f  s%-3s %-28s %s" % r)


if __name__ == "__main__":
    a = sys.argv[1:]
    main(int(a[0]) if a else 4, int(a[1]) if len(a) > 1 else 1, a[2] if len(a) > 2 else "native",
         int(a[3]) if len(a) > 3 else 32, int(a[4]) if len(a) > 4 else 0, int(a[5]) if len(a) > 5 else 0)
