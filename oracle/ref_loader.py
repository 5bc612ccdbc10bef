"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference function in this container.

The reference's hot path lives in ``/root/reference/backend/app.py:174-250``
(``depth_to_point_cloud``).  ``app.py`` imports ``trimesh``, ``open3d`` and ``laspy`` at module
top (``app.py:14-16``); none of them is used by the hot path and none is installed here, so three
empty ``types.ModuleType`` stubs are put into ``sys.modules`` before the file is executed.

The reference tree does not exist on the GPU box, so nothing that runs there (``-m gpu`` tests,
``smoke()``, ``bench.py``) may import this module.  It is used by ``oracle/make_golden.py`` (which
writes the committed fixtures under ``tests/golden/``) and by CPU tests that skip when the tree is
absent.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("D2PC_REFERENCE_ROOT", "/root/reference")
_REF_APP = os.path.join(REFERENCE_ROOT, "backend", "app.py")
_cached = None


def reference_available() -> bool:
    return os.path.isfile(_REF_APP)


def load_reference_module():
    """Execute the reference's ``backend/app.py`` unmodified and return the module object."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"reference not found at {_REF_APP}")
    for name in ("trimesh", "open3d", "laspy"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location("_d2pc_reference_app", _REF_APP)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod


def reference_depth_to_point_cloud():
    """The reference's own ``depth_to_point_cloud`` (``backend/app.py:174``)."""
    return load_reference_module().depth_to_point_cloud
