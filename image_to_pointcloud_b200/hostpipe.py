"""HostFramePipeline -- end-to-end path for many frames that live in HOST memory.

chunk k:   H2D (stream A)  ->  stats + emit kernels (stream B)  ->  D2H (stream C)
with two device slots, so the copies of chunk k+1 / k-1 overlap the kernels of chunk k.  PCIe is
full duplex, so H2D and D2H overlap each other as well.  The only host synchronisation per chunk
is the wait for the chunk's "any frame needs the exact fallback" flag.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import FrameEngine


class HostFramePipeline:
    def __init__(self, img_h: int, img_w: int, dep_h: Optional[int] = None, dep_w: Optional[int] = None, *,
                 img_c: int = 3, chunk: int = 8, density: str = "high", invert: bool = True,
                 depth_scale: float = 10.0, fov: Optional[float] = None, z_range=None,
                 drop_nonfinite: bool = False, device=None):
        self.engine = FrameEngine(img_h, img_w, dep_h, dep_w, batch=chunk, img_c=img_c, device=device)
        eng = self.engine
        self.cfg = eng.make_config(density=density, invert=invert, depth_scale=depth_scale, fov=fov,
                                   z_range=z_range, drop_nonfinite=drop_nonfinite)
        self.chunk = chunk
        self.n_points = eng.points_per_frame(self.cfg)
        self.masked = bool(self.cfg.use_z_range or self.cfg.drop_nonfinite)
        dev = eng.device
        with torch.cuda.device(dev):
            self.s_h2d, self.s_comp, self.s_d2h = (torch.cuda.Stream(dev) for _ in range(3))
            self.slots = []
            for _ in range(2):
                xyz, rgb = eng.alloc_outputs(self.cfg)
                self.slots.append(dict(
                    depth=torch.empty((chunk, eng.dep_h, eng.dep_w), dtype=torch.float32, device=dev),
                    bgr=(torch.empty((chunk, eng.img_h, eng.img_w, img_c), dtype=torch.uint8, device=dev)
                         if img_c >= 3 else None),
                    xyz=xyz, rgb=rgb,
                    count=torch.zeros(chunk, dtype=torch.int32, device=dev),
                    h2d_done=torch.cuda.Event(), comp_done=torch.cuda.Event(), d2h_done=torch.cuda.Event(),
                ))

    # -- pinned-buffer helpers ------------------------------------------------------------
    def alloc_pinned_inputs(self, n_frames: int):
        eng = self.engine
        depths = torch.empty((n_frames, eng.dep_h, eng.dep_w), dtype=torch.float32, pin_memory=True)
        images = (torch.empty((n_frames, eng.img_h, eng.img_w, eng.img_c), dtype=torch.uint8, pin_memory=True)
                  if eng.img_c >= 3 else None)
        return images, depths

    def alloc_pinned_outputs(self, n_frames: int):
        xyz = torch.empty((n_frames, self.n_points, 3), dtype=torch.float32, pin_memory=True)
        rgb = torch.empty((n_frames, self.n_points, 3), dtype=torch.float32, pin_memory=True)
        counts = torch.zeros(n_frames, dtype=torch.int32, pin_memory=True)
        return xyz, rgb, counts

    # -- the pipeline ---------------------------------------------------------------------
    def run_pinned(self, images: Optional[torch.Tensor], depths: torch.Tensor, out_xyz: torch.Tensor,
                   out_rgb: torch.Tensor, out_counts: torch.Tensor) -> None:
        """All arguments are pinned host tensors ([n, ...]).  Returns when every result is on the
        host.  Frames beyond a multiple of ``chunk`` are handled by re-running the last full window
        (results are idempotent), so any n >= chunk works; n < chunk is padded by repetition."""
        eng, cfg, B = self.engine, self.cfg, self.chunk
        n = depths.shape[0]
        if n == 0:
            return
        starts = list(range(0, max(n - B, 0) + 1, B))
        if n >= B and starts[-1] + B < n:
            starts.append(n - B)
        with torch.cuda.device(eng.device):
            pending: Optional[Tuple[int, int]] = None  # (slot, start) whose compute is enqueued
            for k, start in enumerate(starts if n >= B else [0]):
                slot = self.slots[k % 2]
                cnt = min(B, n - start)
                # the slot's previous results must have left the device before we overwrite them
                slot["d2h_done"].synchronize()
                with torch.cuda.stream(self.s_h2d):
                    slot["depth"][:cnt].copy_(depths[start:start + cnt], non_blocking=True)
                    if slot["bgr"] is not None:
                        slot["bgr"][:cnt].copy_(images[start:start + cnt], non_blocking=True)
                    for r in range(cnt, B):  # n < chunk: pad by repeating frame 0
                        slot["depth"][r].copy_(depths[start], non_blocking=True)
                        if slot["bgr"] is not None:
                            slot["bgr"][r].copy_(images[start], non_blocking=True)
                    slot["h2d_done"].record(self.s_h2d)
                self.s_comp.wait_event(slot["h2d_done"])
                eng.enqueue_stats(cfg, slot["depth"], self.s_comp)
                eng.enqueue_status(cfg, self.s_comp)
                eng.enqueue_emit(cfg, slot["depth"], slot["bgr"], slot["xyz"], slot["rgb"], slot["count"],
                                 None, self.s_comp)
                slot["comp_done"].record(self.s_comp)
                if pending is not None:
                    self._drain(pending, out_xyz, out_rgb, out_counts, n)
                pending = (k % 2, start)
                # the fallback flag of this chunk has to be read before the next chunk reuses the
                # engine's status words: wait for this chunk's kernels (copies keep flowing)
                slot["comp_done"].synchronize()
                if int(eng._any_host[0]) != 0:
                    eng.enqueue_stats_fallback(cfg, slot["depth"], self.s_comp)
                    eng.enqueue_emit(cfg, slot["depth"], slot["bgr"], slot["xyz"], slot["rgb"], slot["count"],
                                     None, self.s_comp)
                    slot["comp_done"].record(self.s_comp)
            if pending is not None:
                self._drain(pending, out_xyz, out_rgb, out_counts, n)
            self.s_d2h.synchronize()

    def _drain(self, pending, out_xyz, out_rgb, out_counts, n):
        idx, start = pending
        slot = self.slots[idx]
        cnt = min(self.chunk, n - start)
        self.s_d2h.wait_event(slot["comp_done"])
        with torch.cuda.stream(self.s_d2h):
            out_xyz[start:start + cnt].copy_(slot["xyz"][:cnt], non_blocking=True)
            out_rgb[start:start + cnt].copy_(slot["rgb"][:cnt], non_blocking=True)
            out_counts[start:start + cnt].copy_(slot["count"][:cnt], non_blocking=True)
            slot["d2h_done"].record(self.s_d2h)

    def bytes_per_frame(self) -> Tuple[int, int]:
        eng = self.engine
        h2d = eng.dep_h * eng.dep_w * 4 + (eng.img_h * eng.img_w * eng.img_c if eng.img_c >= 3 else 0)
        d2h = self.n_points * 24 + 4
        return h2d, d2h

    def run(self, images: Sequence[np.ndarray], depths: Sequence[np.ndarray]) -> List[Tuple[np.ndarray, np.ndarray]]:
        """Arbitrary NumPy frames: staged into pinned memory, then ``run_pinned``."""
        n = len(depths)
        pin_img, pin_dep = self.alloc_pinned_inputs(n)
        for i in range(n):
            pin_dep[i].copy_(torch.from_numpy(np.ascontiguousarray(depths[i], dtype=np.float32)))
            if pin_img is not None:
                pin_img[i].copy_(torch.from_numpy(np.ascontiguousarray(images[i])))
        xyz, rgb, counts = self.alloc_pinned_outputs(n)
        self.run_pinned(pin_img, pin_dep, xyz, rgb, counts)
        xs, cs, ks = xyz.numpy(), rgb.numpy(), counts.numpy()
        return [(xs[i, :ks[i]], cs[i, :ks[i]]) for i in range(n)]
