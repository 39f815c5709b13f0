#!/usr/bin/env python
"""Pack the reference's STL fixtures the mesh tests use into tests/golden/mesh_fixtures.npz (RUN IN THE BUILD CONTAINER,
where /root/reference exists; the GPU box does not have it):

    data/6.5mmD_32mmThread_L130mm.STL      BASELINE.json config 4's titanium screw (7 806 triangles)
    tests/resources/10cmcube.stl           100 mm cube (analytic chord lengths)
    tests/resources/threads.stl            a threaded rod (8 891 triangles)
    tests/resources/suzanne.stl            non-convex closed surface

Stored as float32 triangle soups [n, 3, 3] exactly as ``deepdrr_b200.vol.Mesh.from_stl`` reads them, so the GPU tests can build
``Mesh`` objects without the files, and tests/test_mesh_fixtures.py (CPU) checks the loader against the files where they exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200.vol import Mesh  # noqa: E402

FILES = {"screw": "data/6.5mmD_32mmThread_L130mm.STL", "cube": "tests/resources/10cmcube.stl", "threads": "tests/resources/threads.stl",
         "suzanne": "tests/resources/suzanne.stl"}

if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = {}
    for name, rel in FILES.items():
        m = Mesh.from_stl(os.path.join(ref, rel), material="titanium")
        out[name] = m.triangles.astype(np.float32)
        lo, hi = m.get_bounding_AABB
        print(f"{name}: {len(out[name])} triangles, bounds {np.round(lo, 3)} .. {np.round(hi, 3)}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mesh_fixtures.npz"), **out)
