"""SOR alone on a 4K density-high cloud (8.3 M points), CUDA events: python profiles/sor_4k.py [scene|uniform]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402
from profiles.voxel_sweep import depth_maps  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "scene"
H, W = 2160, 3840
dev = torch.device("cuda", 0)
maps, g = depth_maps(dev, H, W)
bgr = torch.randint(0, 256, (1, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
eng = m.FrameEngine(H, W, batch=1, device=dev)
res = eng.process(eng.make_config(density="high"), maps[kind], bgr)
ts = []
for it in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    p, c, idx, st = m.statistical_outlier_removal(res.xyz[0], res.rgb[0], return_device=True)
    b.record()
    torch.cuda.synchronize()
    ts.append(round(a.elapsed_time(b), 2))
print(kind, "4K high", int(res.count[0]), "->", int(p.shape[0]), "ms per call", ts)
