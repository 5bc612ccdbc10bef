"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/writers.npz from the UNMODIFIED reference writers.

Run in the build container (needs /root/reference):  ``python -m oracle.make_golden_writers``

``save_xyz`` (backend/app.py:379-389) is pure Python and runs here as is; its output file is stored
byte for byte.  ``save_las`` / ``save_ply`` need laspy / Open3D, which are not installed: their
layouts are restated in oracle/d2pc_oracle.py and stay "parity unpinned".
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference_module  # noqa: E402
from tests import cases  # noqa: E402


def main():
    ref = load_reference_module()
    p, c = cases.writer_rows()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            os.makedirs("outputs", exist_ok=True)
            path = ref.save_xyz(p, c, "golden")   # app.py:379
            text = open(path, "rb").read()
        finally:
            os.chdir(cwd)
    out = os.path.join(ROOT, "tests", "golden", "writers.npz")
    np.savez_compressed(out, points=p, colors=c, xyz_text=np.frombuffer(text, dtype=np.uint8),
                        numpy_version=np.__version__)
    print("wrote", out, len(text), "bytes of text for", len(p), "rows")


if __name__ == "__main__":
    main()
