#!/usr/bin/env python
"""Development (GPU box): the C2 march against the number of materials (the hot kernel keeps one running total per material and ray
in registers; from five materials on the two-rays-per-lane instantiation spills a little).  The thorax phantom re-segmented into M
density bands, 1536^2, one / two rays per lane."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms
from deepdrr_b200.vol import Volume

names = ["air", "lung", "soft tissue", "muscle", "blood", "bone", "iron", "copper"]
base = phantoms.thorax_volume((512, 512, 400))
carm = phantoms.MobileCArmGeometry()
poses = phantoms.c2_poses(4, seed=1, carm=carm)
dens = np.asarray(base.data)
for M in [int(a) for a in sys.argv[1:]] or [3, 4, 5, 6, 8]:
    edges = np.quantile(dens[::4, ::4, ::4], np.linspace(0, 1, M + 1)[1:-1])
    labels = np.digitize(dens, edges).astype(np.uint16)
    vol = Volume(dens, ({n: i for i, n in enumerate(names[:M])}, labels), anatomical_from_IJK=base.anatomical_from_IJK)
    with Projector(vol, spectrum="120KV_AL43", step=0.1, neglog=True, camera_intrinsics=carm.camera_intrinsics,
                   source_to_detector_distance=carm.source_to_detector_distance) as p:
        out = []
        for rays in (1, 2):
            p.set_rays_per_lane(rays)
            best = 1e9
            for r in range(2):
                p.project(*poses, max_ray_length=carm.max_ray_length)
                best = min(best, p.last_timing_ms()["march"])
            out.append(best / len(poses))
        print(f"M = {M}: one ray per lane {out[0]:.2f} ms/view, two rays per lane {out[1]:.2f} ms/view", flush=True)
