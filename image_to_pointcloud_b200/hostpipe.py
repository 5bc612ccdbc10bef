"""HostFramePipeline -- end-to-end path for many frames that live in HOST memory.

chunk k:   H2D (stream A)  ->  stats + emit kernels (stream B)  ->  D2H (stream C)
with three device slots (each with its own engine: workspace, status words, fallback flag), so the host only
enqueues: the copies of chunks k+1 / k-1 overlap the kernels of chunk k, and PCIe being full duplex, H2D and
D2H overlap each other as well.  The device -> host stream is the bottleneck (24 B per point against 7 B in):
it never waits for the host -- a chunk's copy-out is enqueued behind its kernels at once, and the "a frame
needs the exact fallback" flag is looked at when the slot is retired (rare: the chunk is then redone through
the exact path and copied out again).  The first chunk is short, so the copy-out stream starts early.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import FrameEngine

N_SLOTS = 3


def plan_chunks(n: int, chunk: int) -> List[Tuple[int, int]]:
    """(start, count) of every chunk of an n-frame run: a short first chunk (the copy-out stream starts after a
    quarter of a chunk's upload instead of a whole one), then full chunks, then the remainder.  Every frame is in
    exactly one chunk, no chunk is longer than ``chunk``."""
    out, s = [], 0
    first = min(n, max(1, chunk // 4)) if n > chunk else min(n, chunk)
    while s < n:
        c = first if s == 0 else min(chunk, n - s)
        out.append((s, c))
        s += c
    return out


class HostFramePipeline:
    def __init__(self, img_h: int, img_w: int, dep_h: Optional[int] = None, dep_w: Optional[int] = None, *,
                 img_c: int = 3, chunk: int = 8, density: str = "high", invert: bool = True,
                 depth_scale: float = 10.0, fov: Optional[float] = None, z_range=None,
                 drop_nonfinite: bool = False, device=None):
        self.engines = [FrameEngine(img_h, img_w, dep_h, dep_w, batch=chunk, img_c=img_c, device=device)
                        for _ in range(N_SLOTS)]
        self.engine = eng = self.engines[0]
        self.cfg = eng.make_config(density=density, invert=invert, depth_scale=depth_scale, fov=fov,
                                   z_range=z_range, drop_nonfinite=drop_nonfinite)
        self.chunk = chunk
        self.n_points = eng.points_per_frame(self.cfg)
        self.masked = bool(self.cfg.use_z_range or self.cfg.drop_nonfinite)
        self.fallback_chunks = 0   # chunks that had to be redone through the exact path (diagnostic)
        dev = eng.device
        with torch.cuda.device(dev):
            self.s_h2d, self.s_comp, self.s_d2h = (torch.cuda.Stream(dev) for _ in range(3))
            self.slots = []
            for e in self.engines:
                xyz, rgb = e.alloc_outputs(self.cfg)
                self.slots.append(dict(
                    eng=e,
                    depth=torch.empty((chunk, e.dep_h, e.dep_w), dtype=torch.float32, device=dev),
                    bgr=(torch.empty((chunk, e.img_h, e.img_w, img_c), dtype=torch.uint8, device=dev)
                         if img_c >= 3 else None),
                    xyz=xyz, rgb=rgb,
                    count=torch.zeros(chunk, dtype=torch.int32, device=dev),
                    h2d_done=torch.cuda.Event(), comp_done=torch.cuda.Event(), d2h_done=torch.cuda.Event(),
                    chunk=None,   # (start, cnt) of the chunk in flight in this slot
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
    def chunks(self, n: int) -> List[Tuple[int, int]]:
        return plan_chunks(n, self.chunk)

    def run_pinned(self, images: Optional[torch.Tensor], depths: torch.Tensor, out_xyz: torch.Tensor,
                   out_rgb: torch.Tensor, out_counts: torch.Tensor) -> None:
        """All arguments are pinned host tensors ([n, ...]).  Returns when every result is on the host.
        A chunk shorter than ``chunk`` frames is padded on the device by repeating its first frame (the padded
        results are not copied out)."""
        cfg, B = self.cfg, self.chunk
        n = depths.shape[0]
        if n == 0:
            return
        outs = (out_xyz, out_rgb, out_counts)
        dev = self.engine.device
        with torch.cuda.device(dev):
            for k, (start, cnt) in enumerate(self.chunks(n)):
                slot = self.slots[k % N_SLOTS]
                eng = slot["eng"]
                self._retire(slot, outs)   # the slot's previous chunk is on the host (and was exact)
                with torch.cuda.stream(self.s_h2d):
                    slot["depth"][:cnt].copy_(depths[start:start + cnt], non_blocking=True)
                    if slot["bgr"] is not None:
                        slot["bgr"][:cnt].copy_(images[start:start + cnt], non_blocking=True)
                    if cnt < B:   # pad on the device
                        slot["depth"][cnt:] = slot["depth"][0]
                        if slot["bgr"] is not None:
                            slot["bgr"][cnt:] = slot["bgr"][0]
                    slot["h2d_done"].record(self.s_h2d)
                self.s_comp.wait_event(slot["h2d_done"])
                eng.enqueue_stats(cfg, slot["depth"], self.s_comp)
                eng.enqueue_status(cfg, self.s_comp)
                eng.enqueue_emit(cfg, slot["depth"], slot["bgr"], slot["xyz"], slot["rgb"], slot["count"],
                                 None, self.s_comp)
                slot["comp_done"].record(self.s_comp)
                slot["chunk"] = (start, cnt)
                self._copy_out(slot, outs)
            for j in range(N_SLOTS):   # oldest first
                self._retire(self.slots[(k + 1 + j) % N_SLOTS], outs)

    def _copy_out(self, slot, outs):
        out_xyz, out_rgb, out_counts = outs
        start, cnt = slot["chunk"]
        self.s_d2h.wait_event(slot["comp_done"])
        with torch.cuda.stream(self.s_d2h):
            out_xyz[start:start + cnt].copy_(slot["xyz"][:cnt], non_blocking=True)
            out_rgb[start:start + cnt].copy_(slot["rgb"][:cnt], non_blocking=True)
            out_counts[start:start + cnt].copy_(slot["count"][:cnt], non_blocking=True)
            slot["d2h_done"].record(self.s_d2h)

    def _retire(self, slot, outs):
        """Host side of a slot's chunk: wait until it is on the host; if one of its frames needed the exact
        fallback (flag written by the chunk's status kernel), redo the chunk through it and copy out again."""
        if slot["chunk"] is None:
            return
        eng, cfg = slot["eng"], self.cfg
        slot["d2h_done"].synchronize()
        if int(eng._any_host[0]) != 0:
            self.fallback_chunks += 1
            eng.enqueue_stats_fallback(cfg, slot["depth"], self.s_comp)
            eng.enqueue_emit(cfg, slot["depth"], slot["bgr"], slot["xyz"], slot["rgb"], slot["count"],
                             None, self.s_comp)
            slot["comp_done"].record(self.s_comp)
            self._copy_out(slot, outs)
            slot["d2h_done"].synchronize()
        slot["chunk"] = None

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


class MultiGpuPipeline:
    """Frames in host memory -> point clouds in host memory on every GPU of the node.

    Frames are independent (the reference function is pure, app.py:174-250), so they are sharded by frame
    (``shard_frames``: contiguous, balanced) over one ``HostFramePipeline`` per device, each driven by its own
    host thread -- no collective, no device-to-device traffic.  The library calls and the pinned copies release
    the GIL, so one process feeds all GPUs; a deployment that prefers one process per GPU (``torchrun``) builds
    ``MultiGpuPipeline(devices=[LOCAL_RANK])`` per rank and passes its own ``shard_frames`` slice instead.

    ``devices``: CUDA device indices (default: all visible).  ``pipeline_factory(device_index)`` builds the
    per-device pipeline (default: ``HostFramePipeline`` with the remaining keyword arguments); anything with
    ``run_pinned`` / ``alloc_pinned_inputs`` / ``alloc_pinned_outputs`` works.
    """

    def __init__(self, img_h: int, img_w: int, dep_h: Optional[int] = None, dep_w: Optional[int] = None, *,
                 devices: Optional[Sequence[int]] = None, pipeline_factory=None, **kw):
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        if len(devices) == 0:
            raise RuntimeError("MultiGpuPipeline needs at least one CUDA device")
        if pipeline_factory is None:
            def pipeline_factory(d):
                return HostFramePipeline(img_h, img_w, dep_h, dep_w, device=torch.device("cuda", d), **kw)
        self.devices = list(devices)
        self.pipes = [pipeline_factory(d) for d in self.devices]

    def alloc_pinned_inputs(self, n_frames: int):
        return self.pipes[0].alloc_pinned_inputs(n_frames)

    def alloc_pinned_outputs(self, n_frames: int):
        return self.pipes[0].alloc_pinned_outputs(n_frames)

    def shards(self, n_frames: int) -> List[range]:
        from .engine import shard_frames
        return [shard_frames(n_frames, len(self.pipes), r) for r in range(len(self.pipes))]

    def run_pinned(self, images, depths, out_xyz, out_rgb, out_counts) -> None:
        """Same contract as ``HostFramePipeline.run_pinned``; shard r of the frames runs on device r.
        The first exception of any shard is re-raised after all threads have finished."""
        import threading
        n = depths.shape[0]
        errors: List[BaseException] = []

        def work(pipe, rng):
            try:
                if len(rng) == 0:
                    return
                a, b = rng.start, rng.stop
                pipe.run_pinned(images[a:b] if images is not None else None, depths[a:b], out_xyz[a:b],
                                out_rgb[a:b], out_counts[a:b])
            except BaseException as e:  # noqa: BLE001  (re-raised on the caller's thread)
                errors.append(e)

        threads = [threading.Thread(target=work, args=(p, r), daemon=True) for p, r in zip(self.pipes, self.shards(n))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    def run(self, images: Sequence[np.ndarray], depths: Sequence[np.ndarray]) -> List[Tuple[np.ndarray, np.ndarray]]:
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
