"""Reference arm of bench.py: the UNMODIFIED reference file, run on the host cores.

``install()`` copies ``/root/reference/backend/app.py`` to the git-ignored ``baseline/_ref/backend/app.py``
(the build container has the tree; the GPU box receives the copy with the repo snapshot).  The reference is a
FastAPI application file, not a package: "installing" it is copying that one file.  ``load()`` executes it
unmodified with empty stand-ins for the three top-level imports the hot path never touches
(``trimesh``, ``open3d``, ``laspy``: app.py:14-16, not installed in this image) and returns its
``depth_to_point_cloud`` (app.py:174-250)."""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref", "backend", "app.py")
REF_SOURCE = os.path.join(os.environ.get("D2PC_REFERENCE_ROOT", "/root/reference"), "backend", "app.py")
_fn = None


def install() -> bool:
    """Copy the reference file next to the bench (no-op when the tree is absent).  True if a copy exists."""
    if os.path.isfile(REF_SOURCE):
        os.makedirs(os.path.dirname(REF_COPY), exist_ok=True)
        shutil.copyfile(REF_SOURCE, REF_COPY)
    return os.path.isfile(REF_COPY)


def available() -> bool:
    return os.path.isfile(REF_COPY)


def load():
    """The reference's own depth_to_point_cloud, from the unmodified copy."""
    global _fn
    if _fn is not None:
        return _fn
    for name in ("trimesh", "open3d", "laspy"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location("_reference_backend_app", REF_COPY)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _fn = mod.depth_to_point_cloud
    return _fn
