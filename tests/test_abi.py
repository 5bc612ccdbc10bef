"""CPU: libd2pc.so builds for sm_100a, loads, and exports every symbol include/d2pc.h declares.
Only host-side entry points are called here (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

import image_to_pointcloud_b200 as m
from image_to_pointcloud_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "d2pc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(d2pc_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    return m.load_library()


def test_header_and_library_agree(lib):
    declared = _declared_symbols()
    assert declared == sorted(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/d2pc.h but not exported"
    assert lib.d2pc_abi_version() == 1


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_struct_layout_matches_header(lib):
    # the ctypes mirror must have the C layout: int32*8, double*4, int32, float*2, int32*3
    assert C.sizeof(m.D2pcConfig) == 8 * 4 + 4 * 8 + 4 + 2 * 4 + 3 * 4 == 88
    assert C.sizeof(m.D2pcFrameParams) == 4 * 8 + 4 * 4 + 4 * 4 + 4 * 4 == 80


def test_workspace_bytes_and_validation(lib):
    cfg = m.D2pcConfig(batch=2, img_h=480, img_w=640, img_c=3, dep_h=518, dep_w=686, step=1, invert=1,
                       depth_scale=10.0, cx=320.0, cy=240.0, f=768.0)
    n = C.c_size_t(0)
    assert lib.d2pc_workspace_bytes(C.byref(cfg), C.byref(n)) == 0
    assert n.value > 0 and n.value % 256 == 0
    one = C.c_size_t(0)
    cfg.batch = 1
    assert lib.d2pc_workspace_bytes(C.byref(cfg), C.byref(one)) == 0 and one.value < n.value
    cfg.step = 3
    assert lib.d2pc_workspace_bytes(C.byref(cfg), C.byref(n)) == 1  # invalid argument
    cfg.step = 1
    cfg.dep_h, cfg.dep_w = 1, 7  # a 1-pixel-high map is a valid input (OpenCV's non-IPP resize arithmetic)
    assert lib.d2pc_workspace_bytes(C.byref(cfg), C.byref(n)) == 0
    assert b"unsupported" in lib.d2pc_error_string(4)
    # NULL workspace / depth pointers are rejected before any launch
    cfg.dep_h, cfg.dep_w = 518, 686
    assert lib.d2pc_stats_enqueue(C.byref(cfg), None, None, 0, None) != 0
    assert lib.d2pc_emit_enqueue(C.byref(cfg), None, None, None, 0, None, None, None, None, None) != 0
    t = C.c_size_t(0)
    assert lib.d2pc_voxel_table_bytes(C.byref(cfg), C.byref(t)) == 0 and t.value > 0


def test_reference_signature_and_errors():
    import inspect
    sig = inspect.signature(m.depth_to_point_cloud)
    names = list(sig.parameters)
    # positional part is the reference's signature (backend/app.py:174-180), same defaults
    assert names[:8] == ["image", "depth", "density", "invert", "depth_scale", "smooth", "smooth_ksize", "fov"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["density"], d["invert"], d["depth_scale"], d["smooth"], d["smooth_ksize"], d["fov"]) == \
        ("medium", True, 10.0, False, 5, None)
    for k in ("z_range", "drop_nonfinite", "voxel_size", "device"):
        assert sig.parameters[k].kind is inspect.Parameter.KEYWORD_ONLY
    import numpy as np
    with pytest.raises(KeyError):  # unknown density -> KeyError before any device work (app.py:226)
        m.depth_to_point_cloud(np.zeros((4, 4, 3), np.uint8), np.zeros((4, 4), np.float32), density="ultra")
    with pytest.raises(TypeError):  # cv2 rounds an interpolated integer map back to integers: refused, not approximated
        m.depth_to_point_cloud(np.zeros((4, 4, 3), np.uint8), np.zeros((3, 5), np.uint16))


def test_shard_frames_partition():
    for n in (0, 1, 7, 1024, 1025):
        for ws in (1, 2, 3, 8):
            parts = [m.shard_frames(n, ws, r) for r in range(ws)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "image_to_pointcloud_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_plain_c_host_compiles_links_and_agrees_with_ctypes(lib, tmp_path):
    """include/d2pc.h is a C header (C99, -pedantic clean); a C host linked against libd2pc.so gets the same
    sizes and codes as the ctypes binding."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "host_check")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_host", "host_check.c"), "-o", exe, "-L", libdir, "-ld2pc",
           "-Wl,-rpath," + libdir]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = dict(line.split(" ", 1) for line in out.strip().splitlines())
    cfg = m.D2pcConfig(batch=2, img_h=480, img_w=640, img_c=3, dep_h=518, dep_w=686, step=1, invert=1,
                       depth_scale=10.0, cx=320.0, cy=240.0, f=768.0)
    n = C.c_size_t(0)
    assert int(got["abi_version"]) == lib.d2pc_abi_version()
    assert int(got["sizeof_config"]) == C.sizeof(m.D2pcConfig)
    assert int(got["sizeof_frame_params"]) == C.sizeof(m.D2pcFrameParams)
    assert lib.d2pc_workspace_bytes(C.byref(cfg), C.byref(n)) == 0 and int(got["workspace_bytes"]) == n.value
    assert lib.d2pc_voxel_table_bytes(C.byref(cfg), C.byref(n)) == 0 and int(got["voxel_table_bytes"]) == n.value
    assert lib.d2pc_smooth_scratch_bytes(C.byref(cfg), C.byref(n)) == 0 and int(got["smooth_bytes"]) == n.value
    assert lib.d2pc_sor_scratch_bytes(100000, C.byref(n)) == 0 and int(got["sor_bytes"]) == n.value
    assert lib.d2pc_xyz_text_scratch_bytes(100000, C.byref(n)) == 0 and int(got["text_bytes"]) == n.value
    assert [got[k] for k in ("workspace_rc", "voxel_table_rc", "smooth_rc", "sor_rc", "text_rc")] == ["0"] * 5
    assert got["bad_step_rc"] == "1" and got["null_rc"] == "1" and got["err1"] == "invalid argument"
