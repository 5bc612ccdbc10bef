"""Everything process_image_pipeline does between the depth network and the job result
(reference backend/app.py:468-559) in ONE device-resident pass:

    depth_to_point_cloud (:468-476) -> refine_point_cloud (:479) -> preview decimation (:495-506)
    -> save_point_cloud (:537) -> generate_gis_metadata bounds (:393-400)

Called as separate drop-ins, those steps move the cloud across PCIe five times (down after the stage, up
and down around the refinement, up again for the writer).  Here the image and the depth map go up once,
every step runs on the rows where the previous one left them, and only what the caller keeps comes
down: the refined arrays (optional), the ~20 000 preview rows, the file bytes and six bounds.
Each step is the same kernel sequence the individual drop-ins use, so the results are identical.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import _lib, writers
from ._lib import check
from .api import _engine_for, _image_channels, _to_host, bounds_dict, check_depth_dtype
from .engine import DENSITY_STEP
from .refine import statistical_outlier_removal


def point_cloud_stage(image: np.ndarray, depth: np.ndarray, *, density: str = "medium", invert: bool = True,
                      depth_scale: float = 10.0, smooth: bool = False, smooth_ksize: int = 5, fov: Optional[float] = None,
                      refine: bool = True, nb_neighbors: int = 20, std_ratio: float = 2.0,
                      output_format: Optional[str] = "ply", filename: Optional[str] = None,
                      max_preview: int = writers.MAX_PREVIEW, return_arrays: bool = True, device=None) -> dict:
    """Returns a dict with ``points`` / ``colors`` (host float32 [M,3], if ``return_arrays``), ``point_count``,
    ``preview_points`` / ``preview_colors`` (nested lists as app.py:505-506), ``bounds`` (app.py:393-400),
    and either ``filepath`` (``filename`` given: the file the reference's writer would have put into
    outputs/<filename>.<ext>) or ``file_bytes`` (its content)."""
    DENSITY_STEP[density]  # KeyError first, like the reference
    lib = _lib.load_library()
    img_h, img_w = image.shape[:2]
    dep_h, dep_w = depth.shape[:2]
    check_depth_dtype(depth, int(img_h), int(img_w))
    img_c = _image_channels(image)
    eng = _engine_for(int(img_h), int(img_w), img_c, int(dep_h), int(dep_w), device)
    dev = eng.device
    cfg = eng.make_config(density=density, invert=invert, depth_scale=float(depth_scale), fov=fov)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        d = torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32).reshape(1, dep_h, dep_w)).to(dev)
        im = torch.from_numpy(np.ascontiguousarray(image).reshape(1, img_h, img_w, -1)).to(dev) if img_c >= 3 else None
        res = eng.process(cfg, d, im, stream=stream, smooth_ksize=(smooth_ksize if smooth else None))
        n = int(res.count.cpu()[0])
        xyz, rgb = res.xyz[0, :n], res.rgb[0, :n]
        if refine and n > 0:
            xyz, rgb, _, _ = statistical_outlier_removal(xyz, rgb, nb_neighbors, std_ratio, return_device=True)
        m = int(xyz.shape[0])
        out = {"point_count": m}
        if return_arrays:
            out["points"], out["colors"] = _to_host(xyz), _to_host(rgb)   # pinned D2H at PCIe speed
        pp, pc = writers.preview_lists(xyz, rgb, max_preview)
        out["preview_points"], out["preview_colors"] = pp, pc
        if m > 0:
            cnt = torch.tensor([m], dtype=torch.int32, device=dev)
            keys = torch.empty(6, dtype=torch.int32, device=dev)
            b6 = torch.empty(6, dtype=torch.float32, device=dev)
            check(lib.d2pc_rows_bounds_enqueue(xyz.data_ptr(), cnt.data_ptr(), m, keys.data_ptr(), b6.data_ptr(),
                                               stream.cuda_stream), "d2pc_rows_bounds_enqueue")
            out["bounds"] = bounds_dict(b6.cpu().numpy())
        fmt = (output_format or "").lower()
        if fmt:
            if fmt == "ply":
                rec = writers.ply_vertex_records(xyz, rgb)
                head, ext = writers.PLY_HEADER.format(n=len(rec)).encode("ascii"), "ply"
            elif fmt in ("las", "laz"):
                rec, offsets, mm = writers.las_point_records(xyz, rgb, 0.01)
                head, ext = writers.las_header(len(rec), 0.01, offsets, mm), "las"
            elif fmt == "xyz":
                rec, head, ext = np.frombuffer(writers.xyz_text(xyz, rgb), dtype=np.uint8), b"", "xyz"
            else:
                raise ValueError(f"Unsupported format: {output_format}")
            out["file_ext"] = ext
            if filename is not None:   # straight from the pinned buffer to the file, no concatenated copy
                Path("outputs").mkdir(exist_ok=True)
                path = f"outputs/{filename}.{ext}"
                with open(path, "wb") as f:
                    f.write(head)
                    f.write(memoryview(rec).cast("B"))
                out["filepath"] = path
            else:
                out["file_bytes"] = head + rec.tobytes()
        return out
