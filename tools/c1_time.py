#!/usr/bin/env python
"""Development check (GPU box): config C1 (128^3 CT, 256^2 detector, 90KV_AL40 -- the reference's small CPU-runnable case):
views per second through the public call for batches of 1 / 8 / 64 / 256 views, next to the reference kernel's time per view."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms, geo
from oracle import ref_gpu

vol = phantoms.c1_volume()
pose0, mrl = phantoms.c1_camera()
rng = np.random.default_rng(3)
poses = []
for i in range(256):
    d = np.array([0.3, 1.0, 0.2]) + rng.normal(0, 0.3, 3)
    poses.append(phantoms.c1_camera(direction=tuple(d))[0])
with Projector(vol, spectrum="90KV_AL40", step=0.1, neglog=True, camera_intrinsics=pose0.intrinsic) as p:
    p.project(*poses[:8], max_ray_length=mrl)
    for n in (1, 8, 64, 256):
        best = 1e9
        for r in range(5):
            t0 = time.perf_counter()
            img = p.project(*poses[:n], max_ray_length=mrl)
            best = min(best, time.perf_counter() - t0)
        tm = p.last_timing_ms()
        print(f"ours: {n:4d} views per call: {best * 1e3 / n:.3f} ms per view end to end ({n / best:.0f} DRRs/s), march {tm['march'] / n:.3f} ms per view", flush=True)
    mats, pr = p.all_materials, list(p.priorities)
    w2i, src, ijk = p._pose_arrays(poses[:4])
if ref_gpu.available():
    from deepdrr_b200.scene import remap_labels
    r = ref_gpu.RefProjector([np.ascontiguousarray(vol.data)], [remap_labels(vol, mats)], len(mats), lineint=True)
    W, H = pose0.intrinsic.sensor_size
    for i in range(3):
        t0 = time.perf_counter()
        r.line_integrals(W, H, 0.1, w2i[i], src[i], ijk[i], mrl, priority=pr)
        print(f"ref view {i}: {(time.perf_counter() - t0) * 1e3:.2f} ms wall per line-integral launch (incl. copies)", flush=True)
