#!/usr/bin/env python
"""Development: time the C2 march with several A/B builds of the library (tools/build_variant.sh), one process per build.

    python tools/variant_bench.py [--views N] [--shares 3,4,5] [--check] build/variants/libdrr_a.so ...

Each child projects the bench's own first poses on one resident 512x512x400 volume and prints the best-of-3 march time per
view for every TEX share, plus (--check) the worst relative deviation of the line integrals of view 0 from the c2 golden
(tests/golden/c2.npz: the reference kernel's own output, every 8th pixel).
"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n_views, shares, check):
    from deepdrr_b200 import Projector, phantoms
    cache = "/dev/shm/thorax_hu_c2.npy"
    carm = phantoms.MobileCArmGeometry()
    v2 = phantoms.thorax_volume((512, 512, 400))
    poses = phantoms.c2_poses(max(n_views, 4), seed=1, carm=carm)[:n_views]
    with Projector(v2, spectrum="120KV_AL43", step=0.1, neglog=True, camera_intrinsics=carm.camera_intrinsics,
                   source_to_detector_distance=carm.source_to_detector_distance, sampler="hybrid") as p:
        for share in shares:
            p.set_hybrid_share(share)
            best = 1e9
            for r in range(3):
                img = p.project(*poses, max_ray_length=carm.max_ray_length)
                best = min(best, p.last_timing_ms()["march"])
            print(f"  share={share} views={n_views}: march {best / n_views:.3f} ms/view  mean {float(img.mean()):.6f}", flush=True)
        if check:
            g = np.load(os.path.join(ROOT, "tests", "golden", "c2.npz"))
            W, H, sub = int(g["W"]), int(g["H"]), int(g["sub"])
            for share in shares:
                p.set_hybrid_share(share)
                arrs = (g["w2i_0"].reshape(1, 9), g["src_0"].reshape(1, -1, 3), g["ijk_0"].reshape(1, -1, 12))
                area = p.project_arrays(*arrs, (W, H), float(g["max_ray_length"]), want="area")[0][:, ::sub, ::sub]
                gl = g["lineint_0"]
                worst = max(float((np.abs(area[m] - gl[m])[gl[m] > 0] / gl[m][gl[m] > 0]).max()) for m in range(area.shape[0]))
                print(f"  share={share}: worst line-integral deviation from the reference kernel {worst:.2e} (tolerance 1e-5)", flush=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--child":
        child(int(args[1]), [int(x) for x in args[2].split(",")], args[3] == "1")
        sys.exit(0)
    n_views, shares, check, libs = 4, "4", "0", []
    i = 0
    while i < len(args):
        if args[i] == "--views": n_views = int(args[i + 1]); i += 2
        elif args[i] == "--shares": shares = args[i + 1]; i += 2
        elif args[i] == "--check": check = "1"; i += 1
        else: libs.append(args[i]); i += 1
    for lib in libs:
        print(f"== {lib}", flush=True)
        env = dict(os.environ, DRR_B200_LIB=os.path.join(ROOT, lib) if not os.path.isabs(lib) else lib)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(n_views), shares, check], env=env, timeout=600)
