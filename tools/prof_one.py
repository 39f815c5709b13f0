#!/usr/bin/env python
"""Development: project a few C2 views with one sampler (for ncu / timing)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms

sampler = sys.argv[1] if len(sys.argv) > 1 else "alu"
n_views = int(sys.argv[2]) if len(sys.argv) > 2 else 1
share = int(sys.argv[3]) if len(sys.argv) > 3 else -1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
shape = (512, 512, 400)
v2 = phantoms.thorax_volume(shape)
carm = phantoms.MobileCArmGeometry()
poses = phantoms.c2_poses(max(n_views, 4), seed=1, carm=carm)[:n_views]
with Projector(v2, spectrum="120KV_AL43", step=0.1, neglog=True, device=None, camera_intrinsics=carm.camera_intrinsics,
               source_to_detector_distance=carm.source_to_detector_distance, sampler=sampler) as p:
    if share >= 0:
        p.set_hybrid_share(share)
    for r in range(reps):
        t = time.time()
        img = p.project(*poses, max_ray_length=carm.max_ray_length)
        dt = time.time() - t
        tm = p.last_timing_ms()
        print(f"{sampler} share={share} views={n_views}: march {tm['march']:.2f} ms ({tm['march']/n_views:.2f}/view) post {tm['spectral_post']:.2f} total {tm['total']:.2f} wall {dt*1e3:.1f} S={p.last_sample_count():.4e} img {img.shape} {float(img.mean()):.4f}", flush=True)
