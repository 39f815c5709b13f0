#!/usr/bin/env python
"""Development check (GPU box): config C3 (CT + two K-wire volumes, 384^2) -- this library vs the reference kernel."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms, geo
from deepdrr_b200.scene import SceneTables
from oracle import ref_gpu

n_views = int(sys.argv[1]) if len(sys.argv) > 1 else 4
small = "--small" in sys.argv
vols = phantoms.c3_scene((128, 128, 100), (3.2, 3.2, 4.0)) if small else phantoms.c3_scene()
poses, sdd = phantoms.cone_poses(n_views)
with Projector(vols, spectrum="120KV_AL43", step=0.1, neglog=False, camera_intrinsics=poses[0].intrinsic, source_to_detector_distance=sdd) as p:
    p.project_line_integrals(*poses[:1])  # warm-up (module load, attribute set-up)
    t0 = time.time()
    area = p.project_line_integrals(*poses)
    area = area.reshape((n_views,) + area.shape[-3:])
    tm = p.last_timing_ms()
    print(f"ours: {n_views} views, march {tm['march']:.1f} ms total = {tm['march']/n_views:.2f} ms/view; wall {time.time()-t0:.2f}s", flush=True)
    img = p.project(*poses)
    mats = p.all_materials
    pr = list(p.priorities)
    mrl = p.max_ray_length
    w2i, src, ijk = p._pose_arrays(poses)
if ref_gpu.available():
    from deepdrr_b200.scene import remap_labels
    dens = [np.ascontiguousarray(v.data) for v in vols]
    labs = [remap_labels(v, mats) for v in vols]
    r = ref_gpu.RefProjector(dens, labs, len(mats), lineint=True)
    W, H = poses[0].intrinsic.sensor_size
    for i in range(min(n_views, 2)):
        t0 = time.time()
        li = r.line_integrals(W, H, 0.1, w2i[i], src[i], ijk[i], mrl, priority=pr)
        dt = (time.time() - t0) / len(mats)
        msg = f"ref view {i}: ~{dt*1e3:.1f} ms per launch |"
        for m in range(len(mats)):
            ok = li[m] > 0
            rel = np.abs(area[i, m] - li[m])[ok] / li[m][ok]
            msg += f" {mats[m]} max rel {rel.max() if rel.size else 0:.2e} eq {np.mean(area[i,m]==li[m]):.3f}"
        print(msg, flush=True)
