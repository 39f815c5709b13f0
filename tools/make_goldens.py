#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own GPU kernel.  RUN ON THE GPU BOX:

    gpurun -- 'python tools/make_goldens.py'      (writes gpurun_out/golden/*.npz; copy to tests/golden/)

The reference's Python package cannot be imported anywhere offline (cupy / killeengeo / pyrender are
absent) and its tests hold no numeric vectors, so the golden vectors that pin the
CPU oracle are outputs of the reference's *unmodified CUDA kernel* (oracle/_ref, built by
oracle/Makefile from /root/reference) on the deterministic phantoms of deepdrr_b200/phantoms.py.
Each file stores the exact kernel inputs (matrices) next to the outputs so the oracle test needs
nothing but the phantom recipe.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import geo, phantoms  # noqa: E402
from deepdrr_b200.scene import SceneTables  # noqa: E402
from oracle.ref_gpu import RefProjector  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)


def run_case(name, volumes, projs, max_ray_length, spectrum, step=0.1, sub=1, priorities=None, timing_only=False):
    st = SceneTables(volumes, spectrum, priorities=priorities)
    dens = [v.data for v in volumes]
    sp = [v.spacing for v in volumes]
    ref = RefProjector(dens, st.labels, st.M, sp)
    ref.set_spectrum(st.energies, st.pdf, st.mu)
    refl = RefProjector(dens, st.labels, st.M, sp, lineint=True)
    W, H = projs[0].intrinsic.sensor_size
    rec = {"W": W, "H": H, "step": np.float32(step), "max_ray_length": np.float32(max_ray_length), "sub": sub,
           "materials": np.array(st.all_materials), "priorities": np.array(st.priorities), "spectrum": np.array(spectrum)}
    ms_all = []
    for i, p in enumerate(projs):
        w2i, src, a = geo.pose_arrays(p, volumes)
        inten, pp, ms = ref.project(W, H, step, w2i, src, a, max_ray_length, st.priorities)
        ms_all.append(ms)
        if timing_only and i > 0:
            continue
        li = refl.line_integrals(W, H, step, w2i, src, a, max_ray_length, st.priorities)
        rec[f"w2i_{i}"] = w2i
        rec[f"src_{i}"] = src
        rec[f"ijk_{i}"] = a
        rec[f"intensity_{i}"] = inten[::sub, ::sub].copy()
        rec[f"pprob_{i}"] = pp[::sub, ::sub].copy()
        rec[f"lineint_{i}"] = li[:, ::sub, ::sub].copy()
    rec["kernel_ms"] = np.array(ms_all, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(f"[golden] {name}: {len(projs)} views {W}x{H}, ref kernel ms: {np.round(ms_all, 3).tolist()}", flush=True)
    ref.close()
    refl.close()


def solid_angle_case():
    """calculate_solid_angle of the reference kernel (project_kernel.cu:14-133, reached with a non-null solid_angle pointer,
    :213-216) for a few cameras: the C1 views, an oblique MobileCArm pose on a non-square sensor, and a strongly off-centre one."""
    v = phantoms.c1_volume(16)
    st = SceneTables([v], "90KV_AL40")
    ref = RefProjector([v.data], st.labels, st.M, [v.spacing])
    ref.set_spectrum(st.energies, st.pdf, st.mu)
    cams = []
    for d in ((0.3, 1.0, 0.2), (0.0, 1.0, 0.0), (1.0, 0.0, 1.0)):
        p, mrl = phantoms.c1_camera(direction=d)
        cams.append((p, mrl))
    carm = phantoms.MobileCArmGeometry(sensor_width=192, sensor_height=160, pixel_size=1.5)
    cams.append((carm.camera_projection(0.5, -0.4, (10.0, -20.0, 5.0)), carm.max_ray_length))
    k = geo.CameraIntrinsicTransform(np.array([[900.0, 0, 20.0], [0, 1100.0, 150.0], [0, 0, 1]]), sensor_height=120, sensor_width=100)
    cams.append((phantoms.look_at_projection((100.0, -700.0, 50.0), (-0.1, 1.0, 0.0), (0, 0, 1), k), 3000.0))
    rec = {}
    for i, (p, mrl) in enumerate(cams):
        W, H = p.intrinsic.sensor_size
        w2i, src, a = geo.pose_arrays(p, [v])
        rec[f"w2i_{i}"], rec[f"W_{i}"], rec[f"H_{i}"] = w2i, W, H
        rec[f"solid_{i}"] = ref.solid_angle(W, H, w2i, src, a, mrl)
        print(f"[golden] solid angle view {i}: {W}x{H}, mean {rec[f'solid_{i}'].mean():.6e}", flush=True)
    ref.close()
    np.savez_compressed(os.path.join(OUT, "solid_angle.npz"), **rec)


def main():
    t0 = time.time()
    if "--solid-only" in sys.argv:
        solid_angle_case()
        return
    solid_angle_case()
    # C1: 128^3 cylinder, 256^2, 90 kV; plus an axis-aligned view (zero direction components)
    v1 = phantoms.c1_volume()
    p1, mrl1 = phantoms.c1_camera()
    p1b, _ = phantoms.c1_camera(direction=(0.0, 1.0, 0.0))
    p1c, _ = phantoms.c1_camera(direction=(1.0, 0.0, 1.0))
    run_case("c1", [v1], [p1, p1b, p1c], mrl1, "90KV_AL40")
    # small thorax, oblique C-arm poses, 120 kV, 3 materials
    carm_s = phantoms.MobileCArmGeometry(sensor_width=384, sensor_height=384, pixel_size=0.776)
    vs = phantoms.thorax_volume((128, 128, 100), (3.2, 3.2, 4.0))
    run_case("thorax_small", [vs], phantoms.c2_poses(3, seed=1, carm=carm_s), carm_s.max_ray_length, "120KV_AL43")
    # two overlapping volumes, explicit + default priorities (multi-volume semantics incl. shared label cache)
    w = phantoms.kwire_volume(length_mm=60.0, spacing=0.25, half_width=6)
    phantoms.place_kwire(w, (-10.0, -20.0, 0.0), (0.3, 1.0, 0.1))
    w2 = phantoms.kwire_volume(length_mm=60.0, spacing=0.25, half_width=6)
    phantoms.place_kwire(w2, (10.0, -20.0, 5.0), (-0.3, 1.0, 0.0))
    poses_mv, mrl_mv = phantoms.cone_poses(2, seed=2, sensor=192, pixel=0.6)
    run_case("multivol3", [vs, w, w2], poses_mv, mrl_mv, "90KV_AL40")
    vs2 = phantoms.thorax_volume((96, 96, 80), (3.2, 3.2, 4.0), seed=3)
    vs2.translate((15.0, -10.0, 20.0))
    run_case("multivol2_sameprio", [vs, vs2], poses_mv, mrl_mv, "60KV_AL35", priorities=[0, 0])
    print(f"[golden] small cases done in {time.time() - t0:.1f}s", flush=True)
    # C2 full size: timing of the reference kernel + subsampled golden of view 0
    if "--no-c2" not in sys.argv:
        t1 = time.time()
        v2 = phantoms.thorax_volume()
        print(f"[golden] C2 phantom built in {time.time() - t1:.1f}s", flush=True)
        carm = phantoms.MobileCArmGeometry()
        run_case("c2", [v2], phantoms.c2_poses(4, seed=1, carm=carm), carm.max_ray_length, "120KV_AL43", sub=8,
                 timing_only=True)
    print(f"[golden] total {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
