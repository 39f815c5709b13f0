#!/usr/bin/env python
"""Development: time the C2 march for several TEX shares of the hybrid sampler on one resident volume.

    python tools/share_sweep.py [n_views] [share ...]      (env DRR_B200_LIB selects an A/B build)
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms

n_views = int(sys.argv[1]) if len(sys.argv) > 1 else 4
shares = [int(a) for a in sys.argv[2:]] or [4, 5, 6]
v2 = phantoms.thorax_volume((512, 512, 400))
carm = phantoms.MobileCArmGeometry()
poses = phantoms.c2_poses(max(n_views, 4), seed=1, carm=carm)[:n_views]
with Projector(v2, spectrum="120KV_AL43", step=0.1, neglog=True, device=None, camera_intrinsics=carm.camera_intrinsics,
               source_to_detector_distance=carm.source_to_detector_distance, sampler="hybrid") as p:
    for share in shares:
        p.set_hybrid_share(share)
        best = 1e9
        for r in range(3):
            img = p.project(*poses, max_ray_length=carm.max_ray_length)
            best = min(best, p.last_timing_ms()["march"])
        print(f"share={share} views={n_views}: march {best / n_views:.3f} ms/view  checksum {float(img.mean()):.6f}", flush=True)
