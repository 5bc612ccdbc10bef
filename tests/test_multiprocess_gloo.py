"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path -- frames are sharded by
frame with no collective on the data path; torch.distributed is only used for the barrier and the
max-over-ranks of the timing (bench.py).  The per-frame work here is the oracle standing in for the
device (there is no GPU in the build container)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from image_to_pointcloud_b200.engine import shard_frames
    from oracle import d2pc_oracle as O
    from tests import cases
    mine = shard_frames(n_frames, world, rank)
    points = 0
    sums = []
    for i in mine:  # frame i is seeded 1000 + i on every rank
        img = cases.make_image(12, 16, 1000 + i)
        dep = cases.make_depth(12, 16, 1000 + i, "uniform")
        p, c = O.depth_to_point_cloud(img, dep, density="high")
        points += len(p)
        sums.append(float(p[:, 2].astype(np.float64).sum()))
    # bench.py's aggregation: barrier, max over ranks of the elapsed time, sum of the work
    dist.barrier()
    t = torch.tensor([1.0 + rank, float(points)], dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, (list(mine), sums))
    if rank == 0:
        np.save(os.path.join(out_dir, "res.npy"), np.array([tmax[0].item(), tsum[1].item()]))
        import json
        json.dump(gathered, open(os.path.join(out_dir, "gathered.json"), "w"))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_frame_sharding(tmp_path):
    n_frames, world = 7, 2
    mp.spawn(_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    res = np.load(tmp_path / "res.npy")
    assert res[0] == 2.0                      # max over ranks, not rank 0's own value
    assert res[1] == n_frames * 12 * 16       # every frame processed exactly once
    import json
    gathered = json.load(open(tmp_path / "gathered.json"))
    frames = [i for part, _ in gathered for i in part]
    assert frames == list(range(n_frames))
    # a frame's result does not depend on which rank (or how many ranks) processed it
    sys.path.insert(0, ROOT)
    from oracle import d2pc_oracle as O
    from tests import cases
    flat = [s for _, sums in gathered for s in sums]
    for i in range(n_frames):
        p, _ = O.depth_to_point_cloud(cases.make_image(12, 16, 1000 + i), cases.make_depth(12, 16, 1000 + i, "uniform"),
                                      density="high")
        assert flat[i] == float(p[:, 2].astype(np.float64).sum())


def test_multi_gpu_pipeline_orchestration_on_cpu():
    """MultiGpuPipeline's host logic without a GPU: per-device pipelines are stand-ins that record what they are
    asked to do; every frame is processed exactly once, on the device shard_frames assigns, concurrently, and an
    exception in one shard reaches the caller."""
    import threading
    sys.path.insert(0, ROOT)
    from image_to_pointcloud_b200.engine import shard_frames
    from image_to_pointcloud_b200.hostpipe import MultiGpuPipeline

    seen = {}
    barrier = threading.Barrier(3, timeout=20)

    class Fake:
        def __init__(self, d):
            self.d = d

        def alloc_pinned_inputs(self, n):
            return None, torch.zeros((n, 2, 2))

        def alloc_pinned_outputs(self, n):
            return torch.zeros((n, 4, 3)), torch.zeros((n, 4, 3)), torch.zeros(n, dtype=torch.int32)

        def run_pinned(self, images, depths, xyz, rgb, counts):
            barrier.wait()   # all three shards are in flight at the same time
            seen[self.d] = depths[:, 0, 0].tolist()
            xyz[:] = depths[:, :1, :1] + 100.0
            counts[:] = 4
            if self.d == "boom":
                raise ValueError("shard failed")

    n = 10
    pipe = MultiGpuPipeline(2, 2, devices=[0, 1, 2], pipeline_factory=Fake)
    depths = torch.arange(n, dtype=torch.float32).reshape(n, 1, 1).expand(n, 2, 2).contiguous()
    xyz, rgb, counts = pipe.alloc_pinned_outputs(n)
    pipe.run_pinned(None, depths, xyz, rgb, counts)
    for r in range(3):
        assert seen[r] == [float(i) for i in shard_frames(n, 3, r)]
    assert xyz[:, 0, 0].tolist() == [100.0 + i for i in range(n)] and counts.tolist() == [4] * n
    barrier.reset()
    bad = MultiGpuPipeline(2, 2, devices=[0, "boom", 2], pipeline_factory=Fake)
    with pytest.raises(ValueError):
        bad.run_pinned(None, depths, xyz, rgb, counts)


def test_host_pipeline_chunk_plan():
    """HostFramePipeline's chunking: every frame in exactly one chunk, in order, none longer than the engine's
    batch, the first one short when there is more than one chunk."""
    sys.path.insert(0, ROOT)
    from image_to_pointcloud_b200.hostpipe import plan_chunks
    for chunk in (1, 2, 3, 4, 8, 16):
        for n in (0, 1, 2, 3, 7, 8, 9, 31, 32, 33, 100):
            plan = plan_chunks(n, chunk)
            flat = [i for s0, c in plan for i in range(s0, s0 + c)]
            assert flat == list(range(n)), (chunk, n)
            assert all(0 < c <= chunk for _, c in plan), (chunk, n)
            if n > chunk:
                assert plan[0][1] == max(1, chunk // 4)
            elif n:
                assert plan == [(0, n)]
