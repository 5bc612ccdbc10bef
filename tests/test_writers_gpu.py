"""GPU (-m gpu): row f3.  Preview rows, XYZ text, LAS and PLY records produced by the serialise kernels
(through the C ABI) against the reference's own save_xyz bytes (golden) and the oracle."""
import os

import numpy as np
import pytest

from oracle import d2pc_oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "writers.npz")


@pytest.fixture(scope="module")
def m():
    import torch
    assert torch.cuda.is_available()
    import image_to_pointcloud_b200 as mod
    mod.load_library()
    return mod


def test_xyz_text_equals_reference_file(m, tmp_path, monkeypatch):
    z = np.load(GOLD)
    p, c = cases.writer_rows()
    assert m.xyz_text(p, c) == z["xyz_text"].tobytes()
    # the drop-in writer: same file under outputs/
    monkeypatch.chdir(tmp_path)
    path = m.save_point_cloud(p, c, "xyz", "job1")
    assert path == "outputs/job1.xyz"
    assert open(path, "rb").read() == z["xyz_text"].tobytes()
    # sizes around the tile boundaries, device tensors in
    import torch
    for n in (1, 2, 255, 256, 257, 1023, 3000):
        assert m.xyz_text(torch.from_numpy(p[:n]).cuda(), torch.from_numpy(c[:n]).cuda()) == O.xyz_text(p[:n], c[:n])
    assert m.xyz_text(p[:0], c[:0]) == b""
    bad = c[:10].copy()
    bad[3, 1] = np.nan
    with pytest.raises(ValueError):
        m.xyz_text(p[:10], bad)


def test_xyz_text_of_a_real_cloud(m):
    rng = np.random.default_rng(60)
    img = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    dep = (rng.random((77, 91)) * 20).astype(np.float32)
    p, c = m.depth_to_point_cloud(img, dep, density="high")
    assert m.xyz_text(p, c) == O.xyz_text(p, c)


def test_preview_rows(m):
    rng = np.random.default_rng(61)
    for n in (0, 5, 19999, 20000, 20001, 39999, 40001, 76800, 307200):
        p = rng.standard_normal((n, 3)).astype(np.float32)
        c = rng.integers(0, 256, (n, 3)).astype(np.float32)
        gp, gc = m.preview_rows(p, c)
        wp, wc = O.preview_rows(p, c)
        assert np.array_equal(gp, wp) and np.array_equal(gc, wc), n
    lp, lc = m.preview_lists(p, c)
    assert lp == wp.astype(float).tolist() and lc == wc.astype(float).tolist()


def test_las_and_ply_records(m, tmp_path, monkeypatch):
    p, c = cases.writer_rows()
    ok = np.isfinite(p).all(axis=1) & (np.abs(p) < 1e6).all(axis=1)
    p, c = np.ascontiguousarray(p[ok]), np.ascontiguousarray(c[ok])
    rec, off, mm = m.las_point_records(p, c)
    want, woff = O.las_records(p, c)
    assert off == woff
    assert rec.tobytes() == want.tobytes()
    assert [int(v) for v in mm] == [int(want[k].min()) for k in "XYZ"] + [int(want[k].max()) for k in "XYZ"]
    with pytest.raises(OverflowError):
        m.las_point_records(np.array([[0, 0, 0], [3e7, 0, 0]], np.float32), c[:2])
    ply = m.ply_vertex_records(p, c)
    assert ply.tobytes() == O.ply_records(p, c).tobytes()
    monkeypatch.chdir(tmp_path)
    path = m.save_point_cloud(p, c, "ply", "job2")
    raw = open(path, "rb").read()
    head = O.PLY_HEADER.format(n=len(p)).encode()
    assert raw[:len(head)] == head and raw[len(head):] == O.ply_records(p, c).tobytes()
    path = m.save_point_cloud(p, c, "las", "job3")
    raw = open(path, "rb").read()
    assert raw[:4] == b"LASF" and len(raw) == 227 + 26 * len(p) and raw[227:] == want.tobytes()
    with pytest.raises(ValueError):
        m.save_point_cloud(p, c, "obj", "job4")


def test_point_cloud_stage_equals_the_separate_drop_ins(m, tmp_path, monkeypatch):
    """The device-resident pipeline (app.py:468-559 in one pass) returns exactly what the individual
    drop-ins return when they are chained through host arrays like the reference chains its functions."""
    rng = np.random.default_rng(62)
    img = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
    dep = cases.make_depth(259, 343, 21, "scene")
    monkeypatch.chdir(tmp_path)
    for fmt in ("ply", "xyz", "las"):
        out = m.point_cloud_stage(img, dep, density="medium", output_format=fmt, filename="job")
        p, c = m.depth_to_point_cloud(img, dep, density="medium")
        p, c = m.refine_point_cloud(p, c)
        assert out["point_count"] == len(p) and np.array_equal(out["points"], p) and np.array_equal(out["colors"], c)
        pp, pc = m.preview_lists(p, c)
        assert out["preview_points"] == pp and out["preview_colors"] == pc
        assert out["bounds"]["minX"] == float(p[:, 0].min()) and out["bounds"]["maxZ"] == float(p[:, 2].max())
        path = m.save_point_cloud(p, c, fmt, "ref")
        got, want = open(out["filepath"], "rb").read(), open(path, "rb").read()
        if fmt == "las":   # the header carries the creation day; bodies must agree
            assert got[227:] == want[227:] and got[:4] == b"LASF"
        else:
            assert got == want
        mem = m.point_cloud_stage(img, dep, density="medium", output_format=fmt)   # no filename: bytes in memory
        assert (mem["file_bytes"][227:] == got[227:]) if fmt == "las" else (mem["file_bytes"] == got)
    # against the oracle chain as well
    with __import__("warnings").catch_warnings():
        __import__("warnings").simplefilter("ignore")
        po, co = O.depth_to_point_cloud(img, dep, density="medium")
    keep, _, _ = O.statistical_outlier_removal(po)
    out = m.point_cloud_stage(img, dep, density="medium", output_format=None)
    assert np.array_equal(out["points"].view(np.uint32), po[keep].view(np.uint32))
    assert "file_bytes" not in out
