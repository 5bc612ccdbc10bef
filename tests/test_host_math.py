"""CPU: the product's arithmetic header (csrc/d2pc_math.h, the code the sm_100a kernels run per
pixel and per frame) compiled for the host and checked against the golden vectors of the
unmodified reference and against the NumPy oracle."""
import hashlib
import warnings

import numpy as np
import pytest

from oracle import d2pc_oracle as O
from tests import cases
from tests.conftest import assert_bits_equal
from tests.hostmath import harness


@pytest.fixture(scope="module")
def hm():
    return harness.load()


def test_exact_division_by_constant(hm):
    # reciprocal + two FMA residual steps must equal the IEEE quotient bit for bit
    for mode in range(4):
        assert hm.hm_div_check(3_000_000, 100 + mode, mode) == 0


def test_resize_matches_oracle(hm):
    rng = np.random.default_rng(12)
    for sh, dh in [((19, 27), (24, 32)), ((61, 47), (24, 32)), ((2, 2), (9, 9)), ((518, 686), (480, 640)),
                   ((37, 53), (1, 1)), ((5, 9), (40, 3)), ((1, 7), (12, 20)), ((7, 1), (12, 20)), ((1, 1), (12, 20)),
                   ((1, 64), (48, 64)), ((37, 1), (3, 64)), ((1, 5), (1, 9)), ((5, 1), (9, 1))]:
        d = (rng.random(sh) * 20).astype(np.float32)
        d.ravel()[rng.choice(d.size, max(1, d.size // 50), replace=False)] = np.inf
        d.ravel()[rng.choice(d.size, max(1, d.size // 50), replace=False)] = np.nan
        out = np.empty(dh, np.float32)
        hm.hm_resize(d.ctypes.data, sh[0], sh[1], out.ctypes.data, dh[0], dh[1])
        assert_bits_equal(out, O.resize_bilinear(d, dh[0], dh[1]), f"resize {sh}->{dh}")


@pytest.mark.parametrize("simple", [False, True])
def test_small_goldens_bit_exact(hm, small_golden, simple):
    """simple=True routes eligible frames through the guard-free straight-line path
    (make_simple_norm / simple_point) that emit_fast_kernel runs; both must match the reference."""
    n_simple = 0
    for name in small_golden.names:
        img, dep, kw, pts, cols = small_golden.case(name)
        if kw.get("smooth"):
            continue  # the blur runs in its own kernels (checked on the GPU and in the oracle tests)
        p, c = harness.run_stage(hm, img, dep, simple=simple, **kw)
        n_simple += hm.hm_last_simple()
        assert_bits_equal(p, pts, f"{name} points")
        assert_bits_equal(c, cols, f"{name} colors")
    assert (n_simple >= 14) if simple else (n_simple == 0)


def test_simple_path_equals_generic_path_exactly(hm):
    """Same bits INCLUDING the sign of zero (no canonicalisation): zeros at the clip bound,
    -0.0 inputs, negative depth_scale, centre column/row."""
    rng = np.random.default_rng(77)
    for t in range(40):
        H, W = int(rng.integers(2, 40)), int(rng.integers(2, 50))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        dep = (rng.standard_normal((H, W)) * rng.choice([1e-3, 1.0, 1e5])).astype(np.float32)
        if t % 2 == 0:
            dep = np.maximum(dep, 0)          # ReLU zeros: p2 == 0.0
        if t % 3 == 0:
            dep[dep == 0] = -0.0
        kw = dict(density="high", invert=bool(t % 4 < 2), depth_scale=float(rng.choice([10.0, -2.5, 1e-3])))
        a = harness.run_stage(hm, img, dep, simple=False, **kw)[0]
        b = harness.run_stage(hm, img, dep, simple=True, **kw)[0]
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), t


@pytest.mark.parametrize("name", ["c1_480p_high", "c1_480p_low_noinv", "c2_1080p_dav2", "c2_1080p_scene_nonfinite"])
def test_large_goldens_sha(hm, large_golden, name):
    g = large_golden["cases"][name]
    img, dep, kw = cases.build_case(cases.LARGE_CASES[name])
    p, c = harness.run_stage(hm, img, dep, simple=True, **kw)
    assert hashlib.sha256(p.tobytes()).hexdigest() == g["points_sha256"]
    assert hashlib.sha256(c.tobytes()).hexdigest() == g["colors_sha256"]


def test_random_against_oracle(hm):
    rng = np.random.default_rng(21)
    for t in range(30):
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 50))
        h, w = (H, W) if rng.random() < 0.3 else (int(rng.integers(2, 60)), int(rng.integers(2, 60)))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        dep = (rng.standard_normal((h, w)) * rng.choice([1.0, 100.0])).astype(np.float32)
        if t % 3 == 0:
            k = int(rng.integers(1, h * w // 2 + 2))
            dep.ravel()[rng.choice(h * w, k, replace=False)] = rng.choice(np.array([np.nan, np.inf, -np.inf], np.float32), k)
        if t % 7 == 0:
            dep = np.round(dep)
        kw = dict(density=str(rng.choice(["low", "medium", "high"])), invert=bool(rng.random() < 0.5),
                  depth_scale=float(rng.choice([10.0, 1.0, -3.0])))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            po, co = O.depth_to_point_cloud(img, dep, **kw)
        p, c = harness.run_stage(hm, img, dep, **kw)
        assert_bits_equal(p, po, f"case {t} {kw}")
        assert_bits_equal(c, co, f"case {t}")


def test_mask_interval_equals_z_range(hm):
    """ax-1: the depth-space interval found by bisection must select exactly the values whose
    emitted float32 z lies in [z_min, z_max] -- checked on dense value sets that include the
    interval ends and their float32 neighbours."""
    import ctypes as C
    rng = np.random.default_rng(55)
    for t in range(60):
        p2 = float(rng.uniform(-5, 5))
        p98 = p2 + float(np.exp(rng.uniform(-8, 4)))
        invert = int(t % 2)
        scale = float(rng.choice([10.0, 1.0, 37.5, 1e-3]))
        zr = np.sort(rng.uniform(-0.1, 1.1, 2)) * scale
        if t % 7 == 0:
            zr = np.array([0.0, scale])       # exactly the clip ends
        d = rng.uniform(p2 - 1.0, p98 + 1.0, 20000).astype(np.float32)
        lo, hi = C.c_float(0), C.c_float(0)
        bad = hm.hm_mask_check(p2, p98, invert, scale, float(np.float32(zr[0])), float(np.float32(zr[1])),
                               d.ctypes.data, d.size, C.byref(lo), C.byref(hi))
        assert bad == 0, (t, p2, p98, invert, scale, zr, lo.value, hi.value)
        if lo.value <= hi.value:  # probe the ends and their neighbours
            ends = np.array([lo.value, hi.value], np.float32)
            probe = np.concatenate([ends, np.nextafter(ends, np.float32(-np.inf)), np.nextafter(ends, np.float32(np.inf))])
            probe = probe[np.isfinite(probe)]
            bad = hm.hm_mask_check(p2, p98, invert, scale, float(np.float32(zr[0])), float(np.float32(zr[1])),
                                   probe.ctypes.data, probe.size, C.byref(lo), C.byref(hi))
            assert bad == 0, (t, "ends")
