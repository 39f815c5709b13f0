#!/usr/bin/env python
"""Development: where the end-to-end time of Projector.project goes on C2 (8 views per call)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms, geo

carm = phantoms.MobileCArmGeometry()
vol = phantoms.thorax_volume((512, 512, 400))
poses = phantoms.c2_poses(64, seed=1, carm=carm)
with Projector(vol, spectrum="120KV_AL43", step=0.1, neglog=True, camera_intrinsics=carm.camera_intrinsics,
               source_to_detector_distance=carm.source_to_detector_distance) as p:
    for pipe in (0, 1, 2, 4, 1, 0):
        p.set_pipeline(pipe)
        for s in range(2):
            p.project(*poses[:8], max_ray_length=carm.max_ray_length)
        walls, tot, mar, post, prep = [], [], [], [], []
        for s in range(6):
            batch = poses[8 * s:8 * s + 8]
            t0 = time.perf_counter()
            arr = geo.pose_arrays_batch(batch, [vol])
            t1 = time.perf_counter()
            img = p.project(*batch, max_ray_length=carm.max_ray_length)
            t2 = time.perf_counter()
            tm = p.last_timing_ms()
            walls.append((t2 - t1) * 1e3); prep.append((t1 - t0) * 1e3); tot.append(tm["total"]); mar.append(tm["march"]); post.append(tm["spectral_post"])
        print(f"pipeline={pipe}: wall/call {np.mean(walls):.2f} ms, library total {np.mean(tot):.2f}, march {np.mean(mar):.2f}, post {np.mean(post):.2f}, "
              f"pose math alone {np.mean(prep):.2f} ms", flush=True)
