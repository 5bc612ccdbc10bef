"""Sub-batch pipeline (d2pc_path_enqueue) against the two-phase calls: ms per step over sub-batch sizes,
with / without CUDA graph, overlap and L2 hints; outputs compared bit for bit with the two-phase result.
usage: python profiles/path_sweep.py [native|dav2|4k] [batch]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from image_to_pointcloud_b200 import _lib  # noqa: E402

PEAK = 6451.8


def timeit(fn, dev, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / iters


def main(kind="native", B=128, subs=None):
    dev = torch.device("cuda", 0)
    H, W = (2160, 3840) if kind == "4k" else (1080, 1920)
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
    n = eng.points_per_frame(cfg)
    alg = B * (4 * h * w + 3 * n + 24 * n)

    def two_phase():
        eng.enqueue_stats(cfg, depth, s)
        eng.enqueue_status(cfg, s)
        eng.enqueue_emit(cfg, depth, bgr, xyz, rgb, cnt, None, s)
    two_phase()
    torch.cuda.synchronize(dev)
    ref_xyz, ref_rgb = xyz.clone(), rgb.clone()
    out = {"kind": kind, "batch": B, "alg_bytes": alg}
    ms = timeit(two_phase, dev)
    out["two_phase"] = {"ms": round(ms, 4), "frac": round(alg / ms / 1e6 / PEAK, 4)}
    print("two-phase             %.4f ms  frac %.3f" % (ms, alg / ms / 1e6 / PEAK), flush=True)
    combos = subs or [(0, 3), (0, 4), (0, 6), (0, 8), (0, 12)]
    for S, L in combos:
        # S == 0: the index-ordered path kernel with lookahead L frames; S > 0: the stream pipeline
        for name, kw in ((("ordered", dict(flags=_lib.PATH_ORDERED)),) if S == 0 else
                         (("graph", dict(graph=True)), ("nograph", dict(graph=False)))):
            if name == "nograph" and os.environ.get("SWEEP_NOGRAPH", "0") != "1":
                continue
            xyz.zero_(); rgb.zero_()

            def piped():
                eng.enqueue_path(cfg, depth, bgr, xyz, rgb, cnt, None, s, sub_batch=S, lookahead=L, **kw)
            piped()
            torch.cuda.synchronize(dev)
            same = bool(torch.equal(xyz, ref_xyz) and torch.equal(rgb, ref_rgb)) and int(eng._any_host[0]) == 0
            ms = timeit(piped, dev)
            out["S%d_L%d_%s" % (S, L, name)] = {"ms": round(ms, 4), "frac": round(alg / ms / 1e6 / PEAK, 4), "identical": same}
            print("S=%-3d L=%-2d %-10s %.4f ms  frac %.3f  identical %s" % (S, L, name, ms, alg / ms / 1e6 / PEAK, same), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "native"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    subs = [tuple(int(y) for y in x.split(":")) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
    main(kind, B, subs)
