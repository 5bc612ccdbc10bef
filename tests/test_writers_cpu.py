"""CPU: row f3.  The product's formatting header (csrc/d2pc_format.h, the code the serialise kernels run
per row) compiled for the host and checked against the bytes of the UNMODIFIED reference's save_xyz
(tests/golden/writers.npz, written by oracle/make_golden_writers.py), against Python's own float
formatting on random and adversarial values, and against the oracle's LAS / PLY record restatements."""
import os

import numpy as np
import pytest

from oracle import d2pc_oracle as O
from tests import cases
from tests.hostmath import harness

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "writers.npz")


@pytest.fixture(scope="module")
def hm():
    return harness.load()


def _host_xyz_text(hm, p, c):
    p = np.ascontiguousarray(p, np.float32)
    c = np.ascontiguousarray(c, np.float32)
    buf = np.empty(len(p) * 130 + 16, np.uint8)
    n = hm.hm_xyz_text(p.ctypes.data, c.ctypes.data, len(p), buf.ctypes.data, buf.size)
    assert n >= 0, n
    return buf[:n].tobytes()


def test_oracle_xyz_text_equals_reference_file():
    z = np.load(GOLD)
    assert O.xyz_text(z["points"], z["colors"]) == z["xyz_text"].tobytes()


def test_format_header_equals_reference_file(hm):
    z = np.load(GOLD)
    p, c = cases.writer_rows()
    assert np.array_equal(p.view(np.uint32), z["points"].view(np.uint32))  # same rows as the golden run
    assert _host_xyz_text(hm, p, c) == z["xyz_text"].tobytes()


def test_fixed6_against_python_format(hm):
    rng = np.random.default_rng(50)
    vals = [rng.standard_normal(20000).astype(np.float32) * s for s in (1e-6, 1e-3, 1.0, 1e3, 1e7, 1e11)]
    # every float32 of the form odd / 2^k is a potential decimal tie or near-tie
    vals.append(((rng.integers(0, 1 << 20, 20000) * 2 + 1) / 2.0 ** rng.integers(1, 30, 20000)).astype(np.float32))
    vals.append(rng.integers(0, 1 << 32, 20000, dtype=np.uint64).astype(np.uint32).view(np.float32))  # random bit patterns
    v = np.concatenate(vals)
    v = v[np.isfinite(v) & (np.abs(v) < 2.0 ** 44)]
    p = np.zeros((len(v), 3), np.float32)
    p[:, 0] = v
    p[:, 1] = -v
    p[:, 2] = v[::-1]
    c = np.zeros((len(v), 3), np.float32)
    want = "".join(f"{a:.6f} {b:.6f} {d:.6f} 0 0 0\n" for a, b, d in p).encode()
    assert _host_xyz_text(hm, p, c) == want


def test_nonfinite_and_unformattable(hm):
    p = np.array([[np.nan, np.inf, -np.inf]], np.float32)
    c = np.array([[1, 2, 3]], np.float32)
    assert _host_xyz_text(hm, p, c) == O.xyz_text(p, c) == b"nan inf -inf 1 2 3\n"
    big = np.array([[2.0 ** 45, 0, 0]], np.float32)
    buf = np.empty(256, np.uint8)
    assert hm.hm_xyz_text(big.ctypes.data, c.ctypes.data, 1, buf.ctypes.data, 256) == -1
    bad = np.array([[np.nan, 0, 0]], np.float32)  # int(nan) raises in the reference
    assert hm.hm_xyz_text(p.ctypes.data, bad.ctypes.data, 1, buf.ctypes.data, 256) == -1
    neg = np.array([[-3.7, 300.9, -0.5]], np.float32)  # int() truncates toward zero
    assert _host_xyz_text(hm, np.zeros((1, 3), np.float32), neg) == b"0.000000 0.000000 0.000000 -3 300 0\n"


def test_las_and_ply_records_match_oracle(hm):
    p, c = cases.writer_rows()
    ok = np.isfinite(p).all(axis=1) & (np.abs(p) < 1e6).all(axis=1)
    p, c = np.ascontiguousarray(p[ok]), np.ascontiguousarray(c[ok])
    c[5] = [-3.0, 300.0, 127.6]  # clip + truncation
    rec, off = O.las_records(p, c)
    out = np.empty(len(p) * 26, np.uint8)
    offa = np.array(off, np.float64)
    assert hm.hm_las_records(p.ctypes.data, c.ctypes.data, len(p), offa.ctypes.data, 0.01, out.ctypes.data) == 1
    assert out.tobytes() == rec.tobytes()
    # overflow of the scaled int32 is reported
    far = np.array([[0, 0, 0], [3e7, 0, 0]], np.float32)
    assert hm.hm_las_records(far.ctypes.data, c.ctypes.data, 2, np.zeros(3).ctypes.data, 0.01, out.ctypes.data) == 0
    with pytest.raises(OverflowError):
        O.las_records(far, c[:2])
    prec = O.ply_records(p, c)
    pout = np.empty(len(p) * 27, np.uint8)
    hm.hm_ply_records(p.ctypes.data, c.ctypes.data, len(p), pout.ctypes.data)
    assert pout.tobytes() == prec.tobytes()
    # integral colours survive the / 255 * 255 round trip exactly
    allc = np.stack([np.arange(256)] * 3, axis=1).astype(np.float32)
    assert np.array_equal(O.ply_records(np.zeros((256, 3), np.float32), allc)["red"], np.arange(256))


def test_preview_stride(hm):
    for n in (0, 1, 19999, 20000, 20001, 39999, 40000, 76800, 2073600, 8294400):
        pts = np.zeros((n, 3), np.float32)
        pp, _ = O.preview_rows(pts, pts)
        s = hm.hm_preview_stride(n, 20000)
        assert len(pp) == (0 if n == 0 else (n - 1) // s + 1)
        assert len(pp) <= 40000


def test_sor_oracle_against_brute_force():
    """Row f1: the oracle's k-NN (scipy cKDTree) against an O(n^2) NumPy evaluation of the same definition
    (squared distances accumulated axis by axis in float64, self included, ascending sum)."""
    rng = np.random.default_rng(80)
    p = rng.standard_normal((400, 3)).astype(np.float32)
    p[:7] += 9.0                      # outliers
    p[50:60] = p[40:50]               # exact duplicates
    keep, avg, (mean, std, thr) = O.statistical_outlier_removal(p, 20, 2.0)
    q = p.astype(np.float64)
    d2 = (q[:, None, 0] - q[None, :, 0]) ** 2
    d2 = d2 + (q[:, None, 1] - q[None, :, 1]) ** 2
    d2 = d2 + (q[:, None, 2] - q[None, :, 2]) ** 2
    near = np.sqrt(np.sort(d2, axis=1)[:, :20])
    want = np.cumsum(near, axis=1)[:, -1] / 20
    assert np.array_equal(avg, want)
    pos = want > 0
    m = np.cumsum(np.where(pos, want, 0.0))[-1] / len(p)
    s = np.sqrt(np.cumsum(np.where(pos, (want - m) ** 2, 0.0))[-1] / (len(p) - 1))
    assert (mean, std) == (m, s)
    assert np.array_equal(keep, np.nonzero(pos & (want < m + 2.0 * s))[0])
    assert not set(range(7)) & set(keep.tolist())     # the far points are removed
