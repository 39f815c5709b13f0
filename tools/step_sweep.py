#!/usr/bin/env python
"""Development (GPU box): the C2 march against the step length (the lock-step kernels stage the cells of 32 steps at a time: long
steps make long boxes).  Per step length: march ms per view with the lock-step kernel, with the per-ray kernel, and the reference
kernel's time for one view."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms
from oracle import ref_gpu

carm = phantoms.MobileCArmGeometry()
v2 = phantoms.thorax_volume((512, 512, 400))
poses = phantoms.c2_poses(4, seed=1, carm=carm)
ref = None
for step in [float(a) for a in sys.argv[1:]] or [0.1, 0.25, 0.5, 1.0, 2.0]:
    with Projector(v2, spectrum="120KV_AL43", step=step, neglog=True, camera_intrinsics=carm.camera_intrinsics,
                   source_to_detector_distance=carm.source_to_detector_distance) as p:
        res = {}
        for variant in (0, 1):
            p.set_kernel_variant(variant)
            best = 1e9
            for r in range(3):
                p.project(*poses, max_ray_length=carm.max_ray_length)
                best = min(best, p.last_timing_ms()["march"])
            res[variant] = best / len(poses)
        msg = f"step {step:4.2f} mm: lock-step (auto) {res[0]:7.3f} ms/view, per-ray {res[1]:7.3f} ms/view"
        if ref_gpu.available():
            from deepdrr_b200.scene import SceneTables
            if ref is None:
                st = SceneTables([v2], "120KV_AL43")
                ref = ref_gpu.RefProjector([v2.data], st.labels, st.M, [v2.spacing])
                ref.set_spectrum(st.energies, st.pdf, st.mu)
            w2i, src, ijk = p._pose_arrays(poses[:1])
            ms = min(ref.project(carm.sensor_width, carm.sensor_height, step, w2i[0], src[0], ijk[0], carm.max_ray_length, fetch=False)[2] for _ in range(2))
            msg += f"   (reference kernel: {ms:.1f} ms)"
        print(msg, flush=True)
