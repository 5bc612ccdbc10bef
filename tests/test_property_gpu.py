"""GPU (-m gpu): randomised shapes / values (hypothesis) and the full-size BASELINE
configurations [2] (4K, depth-range mask, 5 mm voxels) and [3] (a batch of 1080p frames) checked against the
oracle and through size-independent properties."""
import warnings

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import d2pc_oracle as O
from tests import cases
from tests.conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import torch
    assert torch.cuda.is_available()
    import image_to_pointcloud_b200 as mod
    mod.load_library()
    return mod


def _oracle(img, dep, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return O.depth_to_point_cloud(img, dep, **kw)


DISTS = ["uniform", "normal", "lognormal", "quantised", "constant", "nonfinite", "negative"]


def _depth(rng, h, w, dist):
    if dist == "uniform":
        return (rng.random((h, w)) * 20).astype(np.float32)
    if dist == "normal":
        return (rng.standard_normal((h, w)) * 5).astype(np.float32)
    if dist == "lognormal":
        return np.exp(rng.standard_normal((h, w)) * 3).astype(np.float32)
    if dist == "quantised":
        return np.round(rng.random((h, w)) * 12).astype(np.float32)
    if dist == "constant":
        return np.full((h, w), float(rng.integers(-3, 4)), np.float32)
    if dist == "negative":
        return (-rng.random((h, w)) * 7 - 1).astype(np.float32)
    d = (rng.random((h, w)) * 20).astype(np.float32)
    k = max(1, h * w // 37)
    d.ravel()[rng.choice(h * w, k, replace=False)] = np.array([np.nan, np.inf, -np.inf], np.float32)[np.arange(k) % 3]
    return d


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
@given(H=st.integers(1, 70), W=st.integers(1, 90), h=st.integers(1, 60), w=st.integers(1, 80),
       native=st.booleans(), dens=st.sampled_from(["low", "medium", "high"]), inv=st.booleans(),
       scale=st.sampled_from([10.0, 1.0, 3.7, 250.0]), dist=st.sampled_from(DISTS), chans=st.sampled_from([3, 3, 4, 1]),
       seed=st.integers(0, 2 ** 31 - 1))
def test_random_shapes_and_values(m, H, W, h, w, native, dens, inv, scale, dist, chans, seed):
    rng = np.random.default_rng(seed)
    if native:
        h, w = H, W
    img = rng.integers(0, 256, (H, W) if chans == 1 else (H, W, chans), dtype=np.uint8)
    dep = _depth(rng, h, w, dist)
    kw = dict(density=dens, invert=inv, depth_scale=scale)
    po, co = _oracle(img, dep, **kw)
    p, c = m.depth_to_point_cloud(img, dep, **kw)
    assert_bits_equal(p, po, f"{(H, W, h, w)} {kw} {dist}")
    assert_bits_equal(c, co, "colours")


def test_config2_4k_mask_and_5mm_voxels(m):
    """BASELINE configs[2] at full size: 3840x2160, depth-range mask, 5 mm voxel grid."""
    img, dep, kw = cases.build_case(cases.LARGE_CASES["c3_4k_dav2"])
    po, co = _oracle(img, dep, **kw)
    keep = O.range_mask(po, 0.5, 9.5)
    p, c = m.depth_to_point_cloud(img, dep, z_range=(0.5, 9.5), **kw)
    assert len(p) == int(keep.sum())
    assert_bits_equal(p, po[keep], "4K masked points")
    assert_bits_equal(c, co[keep], "4K masked colours")
    vp, vc, vidx = O.voxel_downsample(po[keep], co[keep], 0.005)
    gp, gc, gidx = m.depth_to_point_cloud(img, dep, z_range=(0.5, 9.5), voxel_size=0.005, return_voxel_index=True, **kw)
    assert len(gp) == len(vp)
    key = (gidx[:, 0].astype(np.int64) << 42) | (gidx[:, 1].astype(np.int64) << 21) | gidx[:, 2]
    order = np.argsort(key)
    assert np.array_equal(gidx[order], vidx), "voxel indices must be bit-exact"
    np.testing.assert_allclose(gp[order], vp, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gc[order], vc, rtol=1e-6, atol=0)
    # properties: every voxel mean lies inside its voxel; the counts add up to the kept points
    vmin = po[keep].min(axis=0).astype(np.float64) - 0.0025
    lo = vmin + gidx * 0.005
    assert np.all(gp >= lo - 1e-6) and np.all(gp <= lo + 0.005 + 1e-6)


def test_config3_batch_of_1080p_frames(m):
    """BASELINE configs[3] (frame batches): 6 different 1080p frames in one engine call; each frame equals
    the single-frame oracle; colours are an exact gather; the call is idempotent."""
    import torch
    B, H, W = 6, 1080, 1920
    imgs = [cases.make_image(H, W, 1000 + i) for i in range(B)]
    deps = [cases.make_depth(H, W, 1000 + i, "uniform" if i % 2 == 0 else "scene") for i in range(B)]
    eng = m.FrameEngine(H, W, batch=B)
    cfg = eng.make_config(density="high")
    d = torch.from_numpy(np.stack(deps)).cuda()
    im = torch.from_numpy(np.stack(imgs)).cuda()
    res = eng.process(cfg, d, im)
    xyz = res.xyz.cpu().numpy()
    rgb = res.rgb.cpu().numpy()
    for b in range(B):
        assert int(res.count[b]) == H * W
        assert np.array_equal(rgb[b], imgs[b][:, :, ::-1].reshape(-1, 3).astype(np.float32))
        assert xyz[b, :, 2].min() >= 0.0 and xyz[b, :, 2].max() <= 10.0
    for b in (0, 3):
        po, co = _oracle(imgs[b], deps[b], density="high")
        assert_bits_equal(xyz[b], po, f"frame {b}")
    res2 = eng.process(cfg, d, im)
    assert torch.equal(res2.xyz, res.xyz) and torch.equal(res2.rgb, res.rgb)


def test_kernels_stay_inside_their_output_slots(m):
    """Canary rows before and after every output buffer must survive (the pool has no compute-sanitizer):
    unmasked / masked / strided / bounds emit on awkward sizes, then voxels and the writer records."""
    import ctypes as C

    import torch
    PAD = 64  # floats on each side (256 B keeps the 16-byte alignment the ABI asks for)
    rng = np.random.default_rng(90)
    for (H, W, h, w, dens, zr) in [(97, 164, 97, 164, "high", None), (97, 164, 40, 70, "high", (1.0, 9.0)),
                                   (64, 1040, 64, 1040, "medium", None), (50, 47, 50, 47, "low", (0.5, 9.5)),
                                   (33, 4, 33, 4, "high", (2.0, 3.0))]:
        B = 3
        eng = m.FrameEngine(H, W, h, w, batch=B)
        cfg = eng.make_config(density=dens, z_range=zr, want_bounds=True)
        n = eng.points_per_frame(cfg)
        d = torch.from_numpy((rng.random((B, h, w)) * 20).astype(np.float32)).cuda()
        im = torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)).cuda()
        bufs = []
        for _ in range(2):
            flat = torch.full((B * n * 3 + 2 * PAD,), -777.0, dtype=torch.float32, device="cuda")
            bufs.append(flat)
        xyz = bufs[0][PAD:PAD + B * n * 3].view(B, n, 3)
        rgb = bufs[1][PAD:PAD + B * n * 3].view(B, n, 3)
        res = eng.process(cfg, d, im, xyz=xyz, rgb=rgb)
        torch.cuda.synchronize()
        for flat in bufs:
            assert bool((flat[:PAD] == -777.0).all()) and bool((flat[-PAD:] == -777.0).all()), (H, W, dens, zr)
        # rows past count[b] of a masked frame are never written
        for b in range(B):
            k = int(res.count[b])
            assert bool((xyz[b, k:] == -777.0).all()) and bool((rgb[b, k:] == -777.0).all())
            po, co = _oracle(im[b].cpu().numpy(), d[b].cpu().numpy(), density=dens)
            keep = O.range_mask(po, *zr) if zr else np.ones(len(po), bool)
            assert_bits_equal(xyz[b, :k].cpu().numpy(), po[keep], "guarded emit")
    # writer records: exact sizes, canary after the last byte
    lib = m.load_library()
    p, c = cases.writer_rows(n=1000)
    ok = np.isfinite(p).all(axis=1) & (np.abs(p) < 1e6).all(axis=1)
    xyz = torch.from_numpy(np.ascontiguousarray(p[ok])).cuda()
    rgb = torch.from_numpy(np.ascontiguousarray(c[ok])).cuda()
    nrow = xyz.shape[0]
    cnt = torch.tensor([nrow], dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for rec_bytes, call in ((27, "ply"), (26, "las")):
        buf = torch.full((nrow * rec_bytes + 256,), 0xAB, dtype=torch.uint8, device="cuda")
        if call == "ply":
            assert lib.d2pc_ply_records_enqueue(xyz.data_ptr(), rgb.data_ptr(), cnt.data_ptr(), nrow, buf.data_ptr(), s) == 0
        else:
            bounds = torch.cat([xyz.amin(0), xyz.amax(0)]).contiguous()
            mm = torch.zeros(6, dtype=torch.int32, device="cuda")
            err = torch.zeros(1, dtype=torch.int32, device="cuda")
            assert lib.d2pc_las_records_enqueue(xyz.data_ptr(), rgb.data_ptr(), cnt.data_ptr(), nrow, bounds.data_ptr(),
                                                0.01, buf.data_ptr(), mm.data_ptr(), err.data_ptr(), s) == 0
        torch.cuda.synchronize()
        assert bool((buf[nrow * rec_bytes:] == 0xAB).all()), call
