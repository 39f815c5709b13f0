#!/usr/bin/env python
"""The reference's example (example_projector.py:31-44, README.md:60-68) on a synthetic CT, with the B200 projector.

    python examples/example_projector.py [out.npy]

Needs a CUDA GPU (there is no CPU fallback).  Writes the stack of DRRs as a NumPy file and prints timing.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdrr_b200 import HUVolume, Mesh, Projector, phantoms  # noqa: E402
from deepdrr_b200.device import MobileCArm  # noqa: E402


def main():
    # a chest-like CT in Hounsfield units; density and the air / soft tissue / bone segmentation are made on the GPU.
    # With a real scan: ct = Volume.from_nifti("scan.nii.gz") or Volume.from_nrrd("scan.nrrd")
    shape, spacing = (256, 256, 200), (1.6, 1.6, 2.0)
    hu = phantoms.thorax_hu(shape, spacing)
    a = np.eye(4)
    for ax in range(3):
        a[ax, ax] = spacing[ax]
        a[ax, 3] = -spacing[ax] * (shape[ax] - 1) / 2.0
    ct = HUVolume(hu, anatomical_from_IJK=a)

    # a titanium screw as a mesh (additive: its density is added on top of the CT along each ray)
    verts, faces = phantoms.screw_mesh()
    screw = Mesh(verts, faces, material="titanium", tag="screw")
    phantoms.place_kwire(screw, (-20.0, -40.0, 10.0), (0.2, 1.0, 0.1))

    carm = MobileCArm(sensor_width=768, sensor_height=768, pixel_size=0.388)
    with Projector([ct, screw], device=carm, spectrum="90KV_AL40", neglog=True) as projector:
        carm.reposition(np.zeros(3))                                      # the phantom is centred at the world origin
        t0 = time.perf_counter()
        single = projector()                                             # the C-arm's current pose
        alphas = np.linspace(-30, 30, 13)                                # degrees
        sweep = projector(*carm.camera_projections(alphas, np.zeros_like(alphas)))
        dt = time.perf_counter() - t0
        mask = projector.project_seg(carm.get_camera_projection(), tags=["screw"])[0]
    print(f"1 + {len(sweep)} DRRs of {single.shape} in {dt * 1e3:.1f} ms; screw covers {int((mask > 0).sum())} pixels")
    if len(sys.argv) > 1:
        np.save(sys.argv[1], np.concatenate([single[None], sweep]))


if __name__ == "__main__":
    main()
