#!/usr/bin/env python
"""Tiny scenes through every kernel family, for `compute-sanitizer --tool memcheck python tools/sanitize_small.py` on a GPU box."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, HUVolume, phantoms, geo
from deepdrr_b200.vol import Mesh

ct = phantoms.thorax_volume((40, 36, 30), (9.0, 9.0, 12.0), seed=1)
w = phantoms.kwire_volume(length_mm=40.0, spacing=0.5, half_width=3)
phantoms.place_kwire(w, (-10.0, -20.0, 0.0), (0.3, 1.0, 0.1))
sv, sf = phantoms.screw_mesh(rings_per_mm=0.3, segments=8)
screw = Mesh(sv, sf, material="titanium", tag="screw")
phantoms.place_kwire(screw, (10.0, -30.0, 0.0), (0.1, 1.0, 0.2))
bv, bf = phantoms.icosphere(25.0, 1)
ball = Mesh(bv, bf, material="lung", density=0.3, subtractive=True, layer=1, tag="ball")
poses, sdd = phantoms.cone_poses(3, seed=3, sensor=37, pixel=4.0)   # odd sensor size: partial tiles
k = poses[0].intrinsic
for name, objs, kw in (("single", [ct], {}), ("single+noise+collected", [ct], dict(add_noise=True, collected_energy=True, noise_seed=1)),
                       ("multi", [ct, w], {}), ("multi+mesh", [ct, w, screw, ball], {}), ("mesh only", [screw], {}),
                       ("outside air", [ct], dict(attenuate_outside_volume=True))):
    for sampler in ("hybrid", "alu", "tex"):
        with Projector(objs, spectrum="60KV_AL35", camera_intrinsics=k, source_to_detector_distance=sdd, sampler=sampler, **kw) as p:
            img = p.project(*poses)
            assert np.isfinite(img).all(), name
            if any(isinstance(o, Mesh) for o in objs) and sampler == "hybrid":
                p.project_hits(poses[0], tags=["screw", None]); p.project_travel(poses[0], tags=["screw"]); p.project_seg(poses[0], tags=["screw", "ball"])
    print("ok", name, flush=True)
hu = phantoms.thorax_hu((40, 36, 30), (9.0, 9.0, 12.0))
with Projector(HUVolume(hu, anatomical_from_IJK=ct.anatomical_from_IJK), camera_intrinsics=k, source_to_detector_distance=sdd) as p:
    p.project(*poses)
print("ok hu", flush=True)
from deepdrr_b200.device import MobileCArm
carm = MobileCArm(sensor_width=37, sensor_height=29, pixel_size=4.0)
with Projector(ct, device=carm, scatter_num=20000, neglog=False) as p:
    p.project()
print("ok scatter", flush=True)
