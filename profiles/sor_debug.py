"""Chosen cell size / occupancy of the SOR grid on a 1080p cloud: python profiles/sor_debug.py scene|uniform"""
import ctypes as C
import os
import struct
import sys

import numpy as np
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
res = eng.process(eng.make_config(density="high", want_bounds=True), maps[kind], bgr)
xyz, rgb, n = res.xyz[0].contiguous(), res.rgb[0].contiguous(), H * W
lib = m.load_library()
nb = C.c_size_t(0)
lib.d2pc_sor_scratch_bytes(n, C.byref(nb))
scratch = torch.zeros(nb.value, dtype=torch.uint8, device=dev)
oxyz, orgb = torch.empty_like(xyz), torch.empty_like(rgb)
oidx = torch.empty(n, dtype=torch.int32, device=dev)
ocnt = torch.zeros(1, dtype=torch.int32, device=dev)
stats = torch.zeros(4, dtype=torch.float64, device=dev)
s = torch.cuda.current_stream(dev).cuda_stream
rc = lib.d2pc_sor_enqueue(xyz.data_ptr(), rgb.data_ptr(), res.count.data_ptr(), n, res.bounds[0].data_ptr(), 20, 2.0,
                          scratch.data_ptr(), scratch.numel(), oxyz.data_ptr(), orgb.data_ptr(), oidx.data_ptr(),
                          ocnt.data_ptr(), stats.data_ptr(), s)
torch.cuda.synchronize()
hdr = scratch[:256].cpu().numpy().tobytes()
mn = struct.unpack_from("<3d", hdr, 0); ext = struct.unpack_from("<3d", hdr, 24)
h, slack = struct.unpack_from("<2d", hdr, 48)
th = struct.unpack_from("<8d", hdr, 64); occ = struct.unpack_from("<8I", hdr, 128)
dim = struct.unpack_from("<3i", hdr, 160); nn, chosen = struct.unpack_from("<2I", hdr, 172)
print("rc", rc, "bounds", res.bounds[0].cpu().numpy())
print("mn", mn, "ext", ext)
print("trial h", [round(x, 5) for x in th])
print("trial occupied", occ, "points/occupied", [round(n / o, 2) if o else 0 for o in occ])
print("chosen", chosen, "h", h, "dim", dim, "n", nn, "kept", int(ocnt[0]))
# occupancy histogram of the chosen grid, query weighted
cap = 1
while cap < 2 * n and cap < 65536: cap <<= 1
cap = max(cap, 65536)
while cap < 2 * n: cap <<= 1
off0 = 256 + cap * 8
cell_off = scratch[off0:off0 + (cap + 1) * 4].view(torch.int32).cpu().numpy().astype(np.int64)
cnt = np.diff(cell_off)
cnt = cnt[cnt > 0]
print("occupied", len(cnt), "max cell", cnt.max(), "query-weighted mean occupancy", float((cnt * cnt).sum() / cnt.sum()))
print("cells with > 1000 points:", int((cnt > 1000).sum()), "holding", int(cnt[cnt > 1000].sum()), "points")
