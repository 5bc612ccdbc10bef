"""Builds and binds tests/hostmath/hostmath.cpp (the product's d2pc_math.h compiled for the host)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libd2pc_hostmath.so")
SRC = os.path.join(HERE, "hostmath.cpp")
HDR = os.path.join(HERE, "..", "..", "image_to_pointcloud_b200", "csrc", "d2pc_math.h")
HDR2 = os.path.join(HERE, "..", "..", "image_to_pointcloud_b200", "csrc", "d2pc_format.h")


def build(force=False):
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    newest = max(os.path.getmtime(SRC), os.path.getmtime(HDR), os.path.getmtime(HDR2))
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17",
                        SRC, "-o", SO], check=True)
    return SO


def load():
    lib = C.CDLL(build())
    lib.hm_div_check.restype = C.c_long
    lib.hm_div_check.argtypes = [C.c_long, C.c_ulong, C.c_int]
    lib.hm_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    lib.hm_depth_to_point_cloud.restype = C.c_long
    lib.hm_depth_to_point_cloud.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hm_mask_check.restype = C.c_long
    lib.hm_mask_check.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_float, C.c_float,
                                  C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
    lib.hm_xyz_text.restype = C.c_long
    lib.hm_xyz_text.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
    lib.hm_las_records.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_double, C.c_void_p]
    lib.hm_ply_records.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
    lib.hm_preview_stride.restype = C.c_uint
    lib.hm_preview_stride.argtypes = [C.c_uint, C.c_uint]
    return lib


def run_stage(lib, image, depth, density="medium", invert=True, depth_scale=10.0, fov=None, simple=False):
    lib.hm_set_simple(1 if simple else 0)
    from image_to_pointcloud_b200.engine import DENSITY_STEP, reference_intrinsics
    H, W = image.shape[:2]
    h, w = depth.shape[:2]
    Cn = image.shape[2] if image.ndim == 3 and image.shape[2] >= 3 else 1
    step = DENSITY_STEP[density]
    cx, cy, f = reference_intrinsics(W, H, fov)
    n = (-(-H // step)) * (-(-W // step))
    xyz = np.empty((n, 3), np.float32)
    rgb = np.empty((n, 3), np.float32)
    img = np.ascontiguousarray(image)
    dep = np.ascontiguousarray(depth, dtype=np.float32)
    got = lib.hm_depth_to_point_cloud(img.ctypes.data, H, W, Cn, dep.ctypes.data, h, w, step, int(invert),
                                      float(depth_scale), cx, cy, f, xyz.ctypes.data, rgb.ctypes.data, None)
    assert got == n
    return xyz, rgb
