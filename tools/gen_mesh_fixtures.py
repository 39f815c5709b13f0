#!/usr/bin/env python
"""Pack the reference's STL fixtures the mesh tests use into tests/golden/mesh_fixtures.npz (RUN IN THE BUILD CONTAINER,
where /root/reference exists; the GPU box does not have it):

    data/6.5mmD_32mmThread_L130mm.STL      BASELINE.json config 4's titanium screw (7 806 triangles)
    tests/resources/10cmcube.stl           100 mm cube (analytic chord lengths)
    tests/resources/threads.stl            a threaded rod (8 891 triangles)
    tests/resources/suzanne.stl            non-convex closed surface
    tests/resources/meshmesh_1/*.stl       the body and the four cubes of the reference's test_mesh_mesh_1

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
         "suzanne": "tests/resources/suzanne.stl", "mm1_body": "tests/resources/meshmesh_1/body.stl",
         "mm1_cube1": "tests/resources/meshmesh_1/Cube_001.stl", "mm1_cube2": "tests/resources/meshmesh_1/Cube_002.stl",
         "mm1_cube3": "tests/resources/meshmesh_1/Cube_003.stl", "mm1_cube4": "tests/resources/meshmesh_1/Cube_004.stl"}

if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = {}
    for name, rel in FILES.items():
        m = Mesh.from_stl(os.path.join(ref, rel), material="titanium")
        out[name] = m.triangles.astype(np.float32)
        lo, hi = m.get_bounding_AABB
        print(f"{name}: {len(out[name])} triangles, bounds {np.round(lo, 3)} .. {np.round(hi, 3)}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mesh_fixtures.npz"), **out)
    # the reference's own truth image for its mesh-only test (tests/test_core.py:353-470 -> tests/reference/test_mesh_mesh_1.gif,
    # 20 frames of 400 x 400, 8 bit after neglog): every fourth frame, as the regression target of tests/test_reference_mesh_golden.py
    from PIL import Image

    im = Image.open(os.path.join(ref, "tests", "reference", "test_mesh_mesh_1.gif"))
    keep = [0, 4, 8, 12, 16, 19]
    frames = []
    for i in keep:
        im.seek(i)
        frames.append(np.array(im.convert("L")))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_test_mesh_mesh_1.npz"), frames=np.stack(frames), frame_ids=np.array(keep))
    print("test_mesh_mesh_1.gif:", im.n_frames, "frames, kept", keep)
