"""Per-frame trace of the persistent path kernel (timestamps it leaves in the workspace): when each frame's
scan, selections and emit ran.  usage: python profiles/path_trace.py [batch] [D]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402


def main(B=32, D=6, H=1080, W=1920):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    depth = torch.rand((B, H, W), generator=g, device=dev) * 20
    bgr = torch.randint(0, 256, (B, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
    eng = m.FrameEngine(H, W, batch=B, device=dev)
    cfg = eng.make_config(density="high")
    xyz, rgb = eng.alloc_outputs(cfg)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream(dev)
    for _ in range(3):
        eng.enqueue_path(cfg, depth, bgr, xyz, rgb, cnt, None, s, lookahead=D, flags=8)
    torch.cuda.synchronize()
    # locate the scheduler words: the workspace layout ends [.. taps | sched | resized]; native depth: sched is last
    ws = eng.workspace.cpu().numpy()
    off = C.c_size_t(0)
    eng.lib.d2pc_path_trace_offset(C.byref(cfg), C.byref(off))
    tr = ws[off.value:off.value + B * 64].view(np.uint64).reshape(B, 8).astype(np.int64)
    t0 = tr[:, 0].min()
    print("frame  scan0  scanN | sel0: start end | sel1: start end | emit0  emitN   (us from the first scan tile)")
    for f in range(B):
        r = (tr[f] - t0) / 1e3
        print("%4d %7.1f %7.1f | %7.1f %7.1f | %7.1f %7.1f | %7.1f %7.1f" % (f, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]))
    print("total %.1f us, %.2f us per frame" % ((tr[:, 7].max() - t0) / 1e3, (tr[:, 7].max() - t0) / 1e3 / B))


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)
