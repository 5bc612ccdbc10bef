"""Row f3 timings on one 1080p cloud (2 073 600 rows, device-resident): kernels only, CUDA events.
    python profiles/writers_bench.py"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_to_pointcloud_b200 as m  # noqa: E402


def bench(iters=10):
    dev = torch.device("cuda", 0)
    lib = m.load_library()
    H, W = 1080, 1920
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    eng = m.FrameEngine(H, W, batch=1, device=dev)
    cfg = eng.make_config(density="high", want_bounds=True)
    depth = torch.rand((1, H, W), generator=g, device=dev) * 20
    bgr = torch.randint(0, 256, (1, H, W, 3), generator=g, device=dev, dtype=torch.uint8)
    res = eng.process(cfg, depth, bgr)
    xyz, rgb, cnt, bounds = res.xyz[0], res.rgb[0], res.count, res.bounds[0]
    n = H * W
    s = torch.cuda.current_stream(dev).cuda_stream
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    out = {}
    nb = C.c_size_t(0)
    lib.d2pc_xyz_text_scratch_bytes(n, C.byref(nb))
    scratch = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    text = torch.empty(n * 64, dtype=torch.uint8, device=dev)

    def xyz_fn():
        lib.d2pc_xyz_text_measure_enqueue(xyz.data_ptr(), rgb.data_ptr(), cnt.data_ptr(), n, scratch.data_ptr(),
                                          scratch.numel(), total.data_ptr(), err.data_ptr(), s)
        lib.d2pc_xyz_text_write_enqueue(xyz.data_ptr(), rgb.data_ptr(), cnt.data_ptr(), n, scratch.data_ptr(),
                                        scratch.numel(), total.data_ptr(), err.data_ptr(), text.data_ptr(),
                                        text.numel(), s)
    ms = timeit(xyz_fn)
    tb = int(total[0])
    out["xyz_text"] = {"ms": round(ms, 4), "text_bytes": tb, "mrows_per_s": round(n / ms / 1e3, 1),
                       "alg_gbs": round((2 * 24 * n + tb) / ms / 1e6, 1)}
    rec = torch.empty(n * 27 + 16, dtype=torch.uint8, device=dev)
    mm = torch.zeros(6, dtype=torch.int32, device=dev)
    ms = timeit(lambda: lib.d2pc_las_records_enqueue(xyz.data_ptr(), rgb.data_ptr(), cnt.data_ptr(), n, bounds.data_ptr(),
                                                     0.01, rec.data_ptr(), mm.data_ptr(), err.data_ptr(), s))
    out["las_records"] = {"ms": round(ms, 4), "mrows_per_s": round(n / ms / 1e3, 1), "alg_gbs": round((24 + 26) * n / ms / 1e6, 1)}
    ms = timeit(lambda: lib.d2pc_ply_records_enqueue(xyz.data_ptr(), rgb.data_ptr(), cnt.data_ptr(), n, rec.data_ptr(), s))
    out["ply_records"] = {"ms": round(ms, 4), "mrows_per_s": round(n / ms / 1e3, 1), "alg_gbs": round((24 + 27) * n / ms / 1e6, 1)}
    oxyz = torch.empty((40001, 3), dtype=torch.float32, device=dev)
    orgb = torch.empty((40001, 3), dtype=torch.float32, device=dev)
    oc = torch.zeros(1, dtype=torch.int32, device=dev)
    ms = timeit(lambda: lib.d2pc_preview_rows_enqueue(xyz.data_ptr(), rgb.data_ptr(), cnt.data_ptr(), 20000, oxyz.data_ptr(),
                                                      orgb.data_ptr(), 40001, oc.data_ptr(), s))
    out["preview_rows"] = {"ms": round(ms, 4), "rows": int(oc[0])}
    return out


if __name__ == "__main__":
    print(json.dumps({"workload": "one 1080p cloud, 2073600 rows, device-resident", "writers": bench()}))
