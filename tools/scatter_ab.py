#!/usr/bin/env python
"""Development (GPU box): two builds of the scatter kernel on the C5 scene -- same tallies and counters, bit for bit?  How fast?

    python tools/scatter_ab.py [--photons N] libA.so libB.so ...      (one child process per library)
"""
import hashlib, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n):
    from deepdrr_b200 import Projector, phantoms, scatter
    volume = phantoms.thorax_volume((512, 512, 400))
    poses, sdd = phantoms.cone_poses(2, seed=6)

    class Dev:
        source_to_detector_distance = sdd
        camera_intrinsics = poses[0].intrinsic
        detector_height = detector_width = 384 * 0.3

        def get_camera_projection(self):
            return poses[0]

    with Projector(volume, device=Dev(), spectrum="120KV_AL43", step=0.1, neglog=False, scatter_num=n, coefficient_records=False) as p:
        scatter.simulate(p, poses[0], 100000, seed=1)
        for k in range(2):
            tally, counters = scatter.simulate(p, poses[k], n, seed=k)
            ms = p.last_timing_ms()["march"]
            print(f"  view {k}: {n} photons in {ms:.1f} ms = {n / ms * 1e3:.3e} photons/s  tally sha1 {hashlib.sha1(tally.tobytes()).hexdigest()[:12]} "
                  f"sum {int(tally.sum())} counters sha1 {hashlib.sha1(counters.tobytes()).hexdigest()[:12]} n_ray {counters[6]:.0f} n_co {counters[7]:.0f}", flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    if a and a[0] == "--child":
        child(int(a[1])); sys.exit(0)
    n = 20_000_000
    if a and a[0] == "--photons":
        n = int(float(a[1])); a = a[2:]
    for lib in a:
        print("==", lib, flush=True)
        env = dict(os.environ, DRR_B200_LIB=lib if os.path.isabs(lib) else os.path.join(ROOT, lib))
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(n)], env=env, timeout=900)
