import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.getcwd())
import image_to_pointcloud_b200 as m
from oracle import d2pc_oracle as O
from tests import cases
os.makedirs("/tmp/bc/outputs", exist_ok=True); os.chdir("/tmp/bc")
img = cases.make_image(2160, 3840, 2); dep = cases.make_depth(518, 924, 2, "scene")
t0 = time.time(); out = m.point_cloud_stage(img, dep, density="high", output_format="ply", filename="big"); t1 = time.time()
print("4K high stage+SOR+PLY:", out["point_count"], "points", round((t1 - t0) * 1e3, 1), "ms", os.path.getsize(out["filepath"]), "bytes")
p, c = m.depth_to_point_cloud(img, dep, density="high")
txt = m.xyz_text(p[:3000000], c[:3000000]); print("xyz text bytes", len(txt), txt[:60])
# SOR at 4K medium (2 M points) against the oracle
pm, cm = m.depth_to_point_cloud(img, dep, density="medium")
t0 = time.time(); keep, avg, st = O.statistical_outlier_removal(pm); t1 = time.time()
gp, gc, gidx, gst = m.statistical_outlier_removal(pm, cm)
diff = set(keep.tolist()) ^ set(gidx.tolist())
print("4K medium SOR: oracle", len(keep), "gpu", len(gidx), "differ", len(diff), "thr rel err", abs(gst["threshold"] - st[2]) / st[2], "oracle s", round(t1 - t0, 1))
