#!/usr/bin/env python
"""Randomised parity check (GPU box, needs oracle/_ref): single-volume scenes with arbitrary volume poses and cameras --
sources inside the volume, grazing rays, coarse detectors (per-ray kernel variant), short max_ray_length -- vs the reference's
own kernel, per-pixel line integrals and intensity."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdrr_b200 import Projector, phantoms, geo
from deepdrr_b200.scene import SceneTables
from oracle import ref_gpu

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst_l, worst_i, fails = 0.0, 0.0, 0
t0 = time.time()


sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402  (the scene generator is shared with tests/test_gpu_parity.py)

only = int(sys.argv[3]) if len(sys.argv) > 3 else None   # replay one iteration (same random stream) and print details
for it in range(n_iter):
    sc = cases.random_single_volume_scene(rng, it, build=(only is None or it == only))
    v, st, W, H, pixel, sdd, k, pose, mrl, mode, shape = sc["volume"], sc["tables"], sc["W"], sc["H"], sc["pixel"], sc["sdd"], sc["k"], sc["pose"], sc["mrl"], sc["mode"], sc["shape"]
    sampler = ["hybrid", "tex", "alu"][it % 3]
    if only is not None and it != only:
        continue
    if only is not None:
        w2i, src, ijk = geo.pose_arrays(pose, [v])
        ref = ref_gpu.RefProjector([v.data], st.labels, st.M, lineint=True)
        li = ref.line_integrals(W, H, 0.1, w2i, src, ijk, mrl)
        for share in (0, 4, 5, 6, 7, 8):
            with Projector(v, spectrum="90KV_AL40", neglog=False, camera_intrinsics=k, source_to_detector_distance=sdd, sampler="hybrid") as p:
                p.set_hybrid_share(share)
                a = p.project_line_integrals(pose, max_ray_length=mrl); a = a.reshape(a.shape[-3:])
            rel = np.abs(a[0] - li[0]) / np.maximum(li[0], 1e-30) * (li[0] > 0)
            print(f"share {share}: air rel err at (45,32) {rel[45, 32] if rel.shape[0] > 45 and rel.shape[1] > 32 else -1:.3e}; max {rel.max():.3e}; pixels above 2e-6: {(rel > 2e-6).sum()} of {(li[0] > 0).sum()}", flush=True)
        for smp in ("alu",):
            for variant in (0, 1):
                with Projector(v, spectrum="90KV_AL40", neglog=False, camera_intrinsics=k, source_to_detector_distance=sdd, sampler=smp) as p:
                    p.set_kernel_variant(variant)
                    a = p.project_line_integrals(pose, max_ray_length=mrl); a = a.reshape(a.shape[-3:])
                for m in range(st.M):
                    mask = li[m] > 0
                    rel = np.where(mask, np.abs(a[m] - li[m]) / np.maximum(li[m], 1e-30), 0)
                    idx = np.unravel_index(np.argmax(rel), rel.shape)
                    print(f"sampler {smp} variant {variant} mat {m}: max rel {rel.max():.3e} at {idx}: ours {a[m][idx]:.9e} ref {li[m][idx]:.9e} (other mats there: {[float(li[q][idx]) for q in range(st.M)]})", flush=True)
        sys.exit(0)
    with Projector(v, spectrum="90KV_AL40", neglog=False, camera_intrinsics=k, source_to_detector_distance=sdd, sampler=sampler) as p:
        area = p.project_line_integrals(pose, max_ray_length=mrl)
        area = area.reshape(area.shape[-3:])
        img = p.project(pose, max_ray_length=mrl)
    ref = ref_gpu.RefProjector([v.data], st.labels, st.M, lineint=True)
    refp = ref_gpu.RefProjector([v.data], st.labels, st.M)
    refp.set_spectrum(st.energies, st.pdf, st.mu)
    w2i, src, ijk = geo.pose_arrays(pose, [v])
    li = ref.line_integrals(W, H, 0.1, w2i, src, ijk, mrl)
    ri, _, _ = refp.project(W, H, 0.1, w2i, src, ijk, mrl)
    for m in range(st.M):
        mask = li[m] > 0
        leak = np.any(area[m][~mask] != 0)
        diff = np.abs(area[m] - li[m])[mask]
        err = float((diff / li[m][mask]).max()) if mask.any() else 0.0
        worst_l = max(worst_l, err)
        if err > 1e-5 or leak:
            fails += 1
            print(f"FAIL it={it} mode={mode} sampler={sampler} mat={m} err={err:.3e} leak={leak} W={W} H={H} pixel={pixel} shape={shape}", flush=True)
    ei = float((np.abs(img - ri) / np.maximum(np.abs(ri), 1e-30)).max())
    worst_i = max(worst_i, ei)
    if ei > 1e-4:
        fails += 1
        print(f"FAIL it={it} mode={mode} sampler={sampler} intensity err={ei:.3e}", flush=True)
    ref.close(); refp.close()
print(f"{n_iter} scenes, worst relative error: line integrals {worst_l:.3e}, intensity {worst_i:.3e}; failures {fails}; {time.time() - t0:.1f} s", flush=True)
sys.exit(1 if fails else 0)
