#!/usr/bin/env python
"""Development: reproduce the reference's mesh-only test (tests/test_core.py:353-470, test_mesh_mesh_1) with this projector and
compare with the reference's own truth frames (tests/golden/ref_test_mesh_mesh_1.npz, from tests/reference/test_mesh_mesh_1.gif)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_reference_mesh_golden import render_frames  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "ref_test_mesh_mesh_1.npz"))
ids = [int(i) for i in g["frame_ids"]]
imgs = render_frames(ids)
for n, i in enumerate(ids):
    want = g["frames"][n].astype(np.int32)
    for name, cand in (("as is", imgs[n]), ("flip ud", imgs[n][::-1]), ("flip lr", imgs[n][:, ::-1]), ("transpose", imgs[n].T), ("rot180", imgs[n][::-1, ::-1])):
        got = (cand * 255).astype(np.uint8).astype(np.int32)
        d = np.abs(got - want)
        print(f"frame {i:2d} {name:9s}: within 1/255 {np.mean(d <= 1):.4f}, within 2/255 {np.mean(d <= 2):.4f}, mean |d| {d.mean():.3f}, max {d.max()}", flush=True)
    np.save(os.path.join(ROOT, "gpurun_out", f"mm1_frame{i}.npy"), (imgs[n] * 255).astype(np.uint8))
