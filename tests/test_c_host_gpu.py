"""GPU (-m gpu): a plain C host (tests/c_host/host_run.c: include/d2pc.h + the CUDA runtime, no Python, no torch)
runs stats -> emit (and the exact fallback when a frame asks for it) and must produce the oracle's bytes."""
import os
import shutil
import struct
import subprocess
import warnings

import numpy as np
import pytest

from oracle import d2pc_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    from image_to_pointcloud_b200 import _lib
    _lib.load_library()   # builds the in-tree library if it is stale
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    out = str(tmp_path_factory.mktemp("c_host") / "host_run")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "c_host", "host_run.c"), "-o", out, "-L", libdir, "-ld2pc",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + libdir,
           "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


def _run(exe, tmp_path, img, dep, step, invert, scale):
    H, W = img.shape[:2]
    C = 1 if img.ndim == 2 else img.shape[2]
    cx, cy, f = O.intrinsics(W, H, None)
    src, dst = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(src, "wb") as fh:
        fh.write(struct.pack("<7i", H, W, C, dep.shape[0], dep.shape[1], step, int(invert)))
        fh.write(struct.pack("<4d", scale, cx, cy, f))
        fh.write(np.ascontiguousarray(dep, np.float32).tobytes())
        fh.write(np.ascontiguousarray(img).tobytes())
    res = subprocess.run([exe, src, dst], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = open(dst, "rb").read()
    n = struct.unpack_from("<I", raw, 0)[0]
    xyz = np.frombuffer(raw, np.float32, n * 3, 4).reshape(n, 3)
    rgb = np.frombuffer(raw, np.float32, n * 3, 4 + n * 12).reshape(n, 3)
    return xyz, rgb, res.stdout


def test_c_host_runs_the_path(exe, tmp_path):
    rng = np.random.default_rng(90)
    cases = [((96, 160, 3), (96, 160), "high", True, False), ((121, 161, 3), (77, 91), "medium", False, False),
             ((240, 320, 4), (259, 343), "high", True, False), ((64, 80, 3), (64, 80), "low", True, True),
             ((50, 70, 3), (1, 9), "high", True, False)]
    for ishape, dshape, dens, inv, nonfinite in cases:
        img = rng.integers(0, 256, ishape, dtype=np.uint8)
        dep = (rng.random(dshape) * 20).astype(np.float32)
        if nonfinite:
            dep[3, 4] = np.nan
            dep[10, 10] = np.inf
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            po, co = O.depth_to_point_cloud(img, dep, density=dens, invert=inv, depth_scale=10.0)
        xyz, rgb, log = _run(exe, tmp_path, img, dep, {"low": 4, "medium": 2, "high": 1}[dens], inv, 10.0)
        assert xyz.tobytes() == po.tobytes() and rgb.tobytes() == co.tobytes(), (ishape, dshape, dens, log)
        assert ("fallback 1" in log) == nonfinite
