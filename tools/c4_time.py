#!/usr/bin/env python
"""Development check (GPU box): config C4 (CT + screw mesh, 384^2): split lock-step / general path vs the general kernel alone."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms, geo
from deepdrr_b200.vol import Mesh

n_views = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ct = phantoms.thorax_volume()
poses, sdd = phantoms.cone_poses(n_views)
for label, kw in (("additive screw", dict()), ("subtractive screw (carves the CT)", dict(subtractive=True, layer=1))):
    sv, sf = phantoms.screw_mesh()
    screw = Mesh(sv, sf, material="titanium", **kw)
    phantoms.place_kwire(screw, (-20.0, -60.0, 10.0), (0.2, 1.0, 0.1))
    out = {}
    for variant in (0, 1):
        with Projector([ct, screw], spectrum="120KV_AL43", step=0.1, neglog=False, camera_intrinsics=poses[0].intrinsic,
                       source_to_detector_distance=sdd) as p:
            p.set_kernel_variant(variant)
            area = p.project_line_integrals(*poses)
            area = p.project_line_integrals(*poses)
            tm = p.last_timing_ms()
            out[variant] = area
            print(f"{label}: variant {variant}: march {tm['march']/n_views:.2f} ms/view, total {tm['total']/n_views:.2f} ms/view, {len(sf)} triangles", flush=True)
    a, b = out[0], out[1]
    print("   split vs general: max abs diff", float(np.abs(a - b).max()), "bit-equal fraction", float(np.mean(a == b)), flush=True)
