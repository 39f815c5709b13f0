#!/usr/bin/env python
"""Randomised parity check (GPU box, needs oracle/_ref): multi-volume scenes with random placements, priorities, enabled
flags and cameras -- this library (lock-step split path) vs the reference's own kernel, per-pixel line integrals."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from deepdrr_b200 import Projector, phantoms, geo
from deepdrr_b200.scene import SceneTables
from oracle import ref_gpu

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
W = H = 88
worst, fails = 0.0, 0
t0 = time.time()


def rot(rng):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)], [2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)], [2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d]])


for it in range(n_iter):
    kind = it % 2
    if kind == 0:   # CT + two wires (V = 3, M = 4)
        ct = phantoms.thorax_volume((56, 48, 40), (6.0, 7.0, 9.0), seed=int(rng.integers(1 << 30)))
        vols = [ct]
        for _ in range(2):
            w = phantoms.kwire_volume(length_mm=float(rng.uniform(30, 120)), spacing=float(rng.choice([0.25, 0.5, 1.0])), half_width=int(rng.integers(2, 6)))
            phantoms.place_kwire(w, rng.uniform(-60, 60, size=3), rng.normal(size=3))
            vols.append(w)
        pr = [int(x) for x in rng.permutation(3)]
    else:           # two CTs on arbitrary grids (V = 2, M = 3)
        a = phantoms.thorax_volume((48, 48, 36), (7.0, 7.0, 9.0), seed=int(rng.integers(1 << 30)))
        b = phantoms.thorax_volume((40, 44, 30), tuple(rng.uniform(4.0, 9.0, size=3)), seed=int(rng.integers(1 << 30)))
        if rng.random() < 0.5:   # same grid, sub-voxel offset: the shared label cache hits all the time
            b = phantoms.thorax_volume((48, 48, 36), (7.0, 7.0, 9.0), seed=int(rng.integers(1 << 30)))
            b.translate(rng.uniform(-6.0, 6.0, size=3))
        else:
            b.rotate(rot(rng)); b.translate(rng.uniform(-80, 80, size=3))
        vols = [a, b]
        pr = [int(x) for x in rng.permutation(2)]
    en = [1] * len(vols)
    st = SceneTables(vols, "90KV_AL40", priorities=pr)
    carm = phantoms.MobileCArmGeometry(sensor_width=W, sensor_height=H, pixel_size=float(rng.uniform(1.5, 4.0)))
    poses = phantoms.c2_poses(2, seed=int(rng.integers(1 << 30)), carm=carm)
    sampler = ["hybrid", "tex", "alu"][it % 3]
    with Projector(vols, priorities=pr, spectrum="90KV_AL40", neglog=False, camera_intrinsics=carm.camera_intrinsics, sampler=sampler) as p:
        area = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
        area = area.reshape((2,) + area.shape[-3:])
    ref = ref_gpu.RefProjector([v.data for v in vols], st.labels, st.M, lineint=True)
    for n, pose in enumerate(poses):
        w2i, src, ijk = geo.pose_arrays(pose, vols)
        li = ref.line_integrals(W, H, 0.1, w2i, src, ijk, carm.max_ray_length, priority=pr, enabled=en)
        for m in range(st.M):
            mask = li[m] > 0
            leak = np.any(area[n, m][~mask] != 0)
            err = float((np.abs(area[n, m] - li[m])[mask] / li[m][mask]).max()) if mask.any() else 0.0
            worst = max(worst, err)
            if err > 1e-5 or leak:
                fails += 1
                print(f"FAIL it={it} kind={kind} sampler={sampler} view={n} mat={m} err={err:.3e} leak={leak} pr={pr}", flush=True)
    ref.close()
print(f"{n_iter} scenes, worst relative line-integral error {worst:.3e}, failures {fails}, {time.time() - t0:.1f} s", flush=True)
sys.exit(1 if fails else 0)
