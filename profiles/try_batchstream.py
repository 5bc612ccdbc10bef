"""Compare one-stream stepping with BatchStream (stats of batch k+1 overlapped with emit of batch k)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = 20
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
H, W = 1080, 1920
sets = []
for i in range(2):
    depth = torch.rand((B, H, W), generator=g, device=dev) * 20
    bgr = torch.randint(0, 256, (B, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
    sets.append((depth, bgr))
eng = m.FrameEngine(H, W, batch=B, device=dev)
cfg = eng.make_config(density="high")
xyz, rgb = eng.alloc_outputs(cfg)
xyz2, rgb2 = eng.alloc_outputs(cfg)
cnt = torch.zeros(B, dtype=torch.int32, device=dev)
cnt2 = torch.zeros(B, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream(dev)
def one_stream(n):
    for k in range(n):
        d, c = sets[k % 2]
        eng.enqueue_stats(cfg, d, s); eng.enqueue_status(cfg, s); eng.enqueue_emit(cfg, d, c, xyz, rgb, cnt, None, s)
one_stream(3); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(s); one_stream(K); b.record(s); torch.cuda.synchronize()
t1 = a.elapsed_time(b) / K
ref = xyz.clone()
bs = m.BatchStream(H, W, batch=B, device=dev, density="high")
outs = [(xyz, rgb, cnt), (xyz2, rgb2, cnt2)]
def two_stream(n):
    for k in range(n):
        d, c = sets[k % 2]
        o = outs[k % 2]
        bs.submit(d, c, *o)
two_stream(4); bs.finish()
torch.cuda.synchronize()
import time
a.record(bs.s_stats)
t0 = time.perf_counter()
two_stream(K)
bs.finish()
dt = (time.perf_counter() - t0) / K * 1e3
print(f"batch {B}: one stream {t1:.4f} ms/step   BatchStream {dt:.4f} ms/step (wall)   fallback batches {bs.needs_fallback()}")
print("same bits:", bool(torch.equal(outs[(K - 1) % 2][0], ref)) )
