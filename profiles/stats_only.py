"""Statistics kernels alone (sample / scan / select) on a batch of synthetic 1080p frames: CUDA-event timing,
or a short run for `ncu --set full --import-source on -k regex:"sample_kernel|scan_kernel|select_kernel"`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402


def main(batch=128, iters=20, H=1080, W=1920, h=None, w=None):
    dev = torch.device("cuda", 0)
    h, w = h or H, w or W
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    depth = torch.rand((batch, h, w), generator=g, device=dev) * 20
    eng = m.FrameEngine(H, W, h, w, batch=batch, device=dev)
    cfg = eng.make_config(density="high")
    s = torch.cuda.current_stream(dev)
    for _ in range(3):
        eng.enqueue_stats(cfg, depth, s)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        eng.enqueue_stats(cfg, depth, s)
    b.record()
    torch.cuda.synchronize()
    eng.enqueue_status(cfg, s)
    torch.cuda.synchronize()
    print("stats ms per %d frames: %.4f  (any fallback: %d)" % (batch, a.elapsed_time(b) / iters, int(eng._any_host[0])))


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)
