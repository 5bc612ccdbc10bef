"""What the box's PCIe links can do, to judge the end-to-end leg against: plain pinned cudaMemcpyAsync (torch
copy_ on a stream, one call per copy) of the pipeline's own sizes, D2H alone, H2D alone and both directions at once,
on 1 .. N GPUs concurrently (one thread per GPU, all started together).  GB/s per GPU and aggregate.

    python profiles/pcie_ceiling.py [max_gpus]        (run on a multi-GPU box: gpurun --gpus 8)
"""
import json
import sys
import threading
import time

import torch

FRAME_OUT = 1080 * 1920 * 24       # xyz + rgb rows of a 1080p frame
FRAME_IN = 1080 * 1920 * 7         # depth + BGR
CHUNK = 8                          # frames per copy, like HostFramePipeline


def worker(dev, mode, n_copies, barrier, out):
    torch.cuda.set_device(dev)
    d_out = torch.empty(CHUNK * FRAME_OUT, dtype=torch.uint8, device=f"cuda:{dev}")
    d_in = torch.empty(CHUNK * FRAME_IN, dtype=torch.uint8, device=f"cuda:{dev}")
    h_out = torch.empty(CHUNK * FRAME_OUT, dtype=torch.uint8, pin_memory=True)
    h_in = torch.empty(CHUNK * FRAME_IN, dtype=torch.uint8, pin_memory=True)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(n):
        for _ in range(n):
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s1):
                    h_out.copy_(d_out, non_blocking=True)
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s2):
                    d_in.copy_(h_in, non_blocking=True)
        s1.synchronize(); s2.synchronize()
    run(3)
    barrier.wait()
    t0 = time.perf_counter()
    run(n_copies)
    dt = time.perf_counter() - t0
    out[dev] = dt


def main(max_gpus=None):
    n_dev = torch.cuda.device_count()
    max_gpus = min(max_gpus or n_dev, n_dev)
    res = {"chunk_frames": CHUNK, "d2h_bytes_per_copy": CHUNK * FRAME_OUT, "h2d_bytes_per_copy": CHUNK * FRAME_IN, "runs": []}
    n_copies = 20
    g = 1
    while g <= max_gpus:
        for mode in ("d2h", "h2d", "both"):
            out = {}
            barrier = threading.Barrier(g)
            th = [threading.Thread(target=worker, args=(d, mode, n_copies, barrier, out)) for d in range(g)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = max(out.values())
            d2h = CHUNK * FRAME_OUT * n_copies * g / dt / 1e9 if mode in ("d2h", "both") else 0.0
            h2d = CHUNK * FRAME_IN * n_copies * g / dt / 1e9 if mode in ("h2d", "both") else 0.0
            row = {"gpus": g, "mode": mode, "d2h_gbs_aggregate": round(d2h, 1), "h2d_gbs_aggregate": round(h2d, 1),
                   "d2h_gbs_per_gpu": round(d2h / g, 1), "h2d_gbs_per_gpu": round(h2d / g, 1)}
            if mode == "both":   # the pipeline moves 24 B out and 7 B in per point: its ceiling in Mpoints/s
                row["e2e_ceiling_mpoints_s"] = round(min(d2h / 24, h2d / 7) * 1e3, 0) if h2d else None
            res["runs"].append(row)
            print(json.dumps(row), flush=True)
        g *= 2
    print(json.dumps(res))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else None)
