#!/usr/bin/env python
"""Development (GPU box): rays per lane (DRR_TUNE_RAYS_PER_LANE) and lane layout (DRR_TUNE_LANE_QUADS) of the single-volume march
against the ray spacing.

One resident C2 volume, the bench's first poses, the same field of view on detectors of N^2 pixels: march ms per view and
per 10^6 rays with one / two rays per lane and the library's own choice (0), then with 2 x 2 lane groups (1) and 4 x 1 runs (0),
for which the images must be identical.
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms

v2 = phantoms.thorax_volume((512, 512, 400))
for n in [int(a) for a in sys.argv[1:]] or [1536, 1152, 1024, 768, 640, 512]:
    carm = phantoms.MobileCArmGeometry(sensor_width=n, sensor_height=n, pixel_size=0.194 * 1536 / n)
    nv = max(4, min(32, int(8 * (1536 / n) ** 2)))
    poses = phantoms.c2_poses(nv, seed=1, carm=carm)
    with Projector(v2, spectrum="120KV_AL43", step=0.1, neglog=True, camera_intrinsics=carm.camera_intrinsics,
                   source_to_detector_distance=carm.source_to_detector_distance, sampler="hybrid") as p:
        out = {}
        for rays in (1, 2, 0):   # 0 = the library's choice from the ray spacing
            p.set_rays_per_lane(rays)
            best = 1e9
            for r in range(3):
                img = p.project(*poses, max_ray_length=carm.max_ray_length)
                best = min(best, p.last_timing_ms()["march"])
            print(f"  {n}^2 rays_per_lane={rays}: {best / nv:.3f} ms/view, {best / nv / (n * n) * 1e6:.3f} ms per Mray", flush=True)
        for mode in (0, 1, 0, 1):
            p.set_lane_quads(mode)
            best = 1e9
            for r in range(3):
                img = p.project(*poses, max_ray_length=carm.max_ray_length)
                best = min(best, p.last_timing_ms()["march"])
            same = mode not in out or np.array_equal(out[mode], img)
            out[mode] = img.copy()
            print(f"  {n}^2 quads={mode}: {best / nv:.3f} ms/view, {best / nv / (n * n) * 1e6:.3f} ms per Mray  identical={same and (0 not in out or 1 not in out or np.array_equal(out[0], out[1]))}", flush=True)
