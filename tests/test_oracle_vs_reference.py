"""CPU, build container only: the oracle against the UNMODIFIED reference function imported
from /root/reference (backend/app.py:174-250).  Skipped where the reference tree is absent
(e.g. the GPU box); the committed goldens cover that case."""
import warnings

import numpy as np
import pytest

from oracle import d2pc_oracle as O
from oracle.ref_loader import reference_available
from tests.conftest import assert_bits_equal

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="/root/reference not present")]


@pytest.fixture(scope="module")
def ref():
    from oracle.ref_loader import reference_depth_to_point_cloud
    return reference_depth_to_point_cloud()


def _cmp(ref, img, dep, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pr, cr = ref(img, dep, **kw)
        po, co = O.depth_to_point_cloud(img, dep, **kw)
    assert_bits_equal(po, pr, f"points {kw}")
    assert_bits_equal(co, cr, f"colors {kw}")


def test_random_shapes_and_knobs(ref):
    rng = np.random.default_rng(99)
    for t in range(40):
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 50))
        if rng.random() < 0.25:
            h, w = H, W
        else:
            h, w = int(rng.integers(2, 60)), int(rng.integers(2, 60))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        kind = t % 5
        if kind == 0:
            dep = (rng.random((h, w)) * 20).astype(np.float32)
        elif kind == 1:
            dep = (rng.standard_normal((h, w)) * 1e3).astype(np.float32)
        elif kind == 2:
            dep = np.round(rng.random((h, w)) * 6).astype(np.float32)
        elif kind == 3:
            dep = (rng.random((h, w)) * 20).astype(np.float32)
            k = int(rng.integers(1, h * w // 2 + 2))
            dep.ravel()[rng.choice(h * w, k, replace=False)] = rng.choice(
                np.array([np.nan, np.inf, -np.inf], np.float32), k)
        else:
            dep = np.full((h, w), float(rng.integers(-3, 4)), np.float32)
            if rng.random() < 0.5:
                dep.ravel()[rng.integers(0, h * w)] += 2.0
        kw = dict(density=str(rng.choice(["low", "medium", "high"])), invert=bool(rng.random() < 0.5),
                  depth_scale=float(rng.choice([10.0, 1.0, 15.0, 0.37])))
        if rng.random() < 0.2:
            kw["fov"] = float(rng.choice([45.0, 60.0, 90.0]))
        _cmp(ref, img, dep, **kw)


def test_c1_480p(ref):
    from tests import cases
    img, dep, kw = cases.build_case(cases.LARGE_CASES["c1_480p_medium"])
    _cmp(ref, img, dep, **kw)


def test_drop_in_signature_equals_the_reference_signature(ref):
    """The drop-in's positional parameters, their order, kinds and defaults are the reference's own
    (backend/app.py:174-180), compared with inspect on the loaded reference function; everything the drop-in adds
    is keyword-only.  The same for the other rebindable functions."""
    import inspect

    import image_to_pointcloud_b200 as m
    from oracle import ref_loader
    rs, ms = inspect.signature(ref), inspect.signature(m.depth_to_point_cloud)
    rp, mp = list(rs.parameters.values()), list(ms.parameters.values())
    assert [(p.name, p.kind, p.default) for p in mp[:len(rp)]] == [(p.name, p.kind, p.default) for p in rp]
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for p in mp[len(rp):])
    app = ref_loader.load_reference_module() if hasattr(ref_loader, "load_reference_module") else None
    if app is not None:
        for name in ("create_depth_preview", "refine_point_cloud", "save_point_cloud", "save_xyz", "save_las", "save_ply"):
            r = list(inspect.signature(getattr(app, name)).parameters.values())
            g = list(inspect.signature(getattr(m, name)).parameters.values())
            assert [(p.name, p.default) for p in g[:len(r)]] == [(p.name, p.default) for p in r], name
            assert all(p.default is not inspect.Parameter.empty or p.kind is inspect.Parameter.KEYWORD_ONLY
                       for p in g[len(r):]), name
