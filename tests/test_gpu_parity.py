"""Parity tests proper: the CUDA path (through Projector -> ctypes -> libdrr_b200.so) against
  (a) golden outputs of the reference's own CUDA kernel (tests/golden, reference run on a B200),
  (b) the CPU oracle on fresh poses,
  (c) the reference kernel itself, live, when oracle/_ref was shipped with the repo snapshot,
plus size-independent properties at BASELINE.json's full sizes and the reference's edge cases.

Tolerances (north_star): per-material line integrals 1e-5 relative per pixel, intensity 1e-4 relative;
labels / indexing are exercised through exact-zero and bit-equality checks.
"""
import numpy as np
import pytest

import cases
from deepdrr_b200 import Projector, geo, phantoms
from oracle import cpu_oracle

pytestmark = pytest.mark.gpu

LINE_RTOL = 1e-5
INT_RTOL = 1e-4


def _projector(volumes, spectrum, priorities, sampler, W, H, **kw):
    k = geo.CameraIntrinsicTransform.from_sizes((W, H), 1.0, 1000.0)
    return Projector(volumes, priorities=priorities, spectrum=spectrum, step=kw.pop("step", 0.1), neglog=kw.pop("neglog", False),
                     camera_intrinsics=k, sampler=sampler, **kw)


def _compare_with_golden(name, samplers, views=None):
    volumes, spectrum, priorities = cases.scene(name)
    g = cases.golden(name)
    W, H, sub = int(g["W"]), int(g["H"]), int(g["sub"])
    nv = cases.n_views(g)
    ids = [i for i in range(nv) if views is None or i in views]
    w2i = np.stack([g[f"w2i_{i}"] for i in ids])
    src = np.stack([g[f"src_{i}"] for i in ids])
    ijk = np.stack([g[f"ijk_{i}"] for i in ids])
    for sampler in samplers:
        with _projector(volumes, spectrum, priorities, sampler, W, H, step=float(g["step"])) as p:
            area = p.project_arrays(w2i, src, ijk, (W, H), float(g["max_ray_length"]), want="area")
            inten = p.project_arrays(w2i, src, ijk, (W, H), float(g["max_ray_length"]), want="intensity", raw=True)
            assert p.launch_count() > 0
        for n, i in enumerate(ids):
            gl, gi = g[f"lineint_{i}"], g[f"intensity_{i}"]
            a, im = area[n][:, ::sub, ::sub], inten[n][::sub, ::sub]
            for m in range(a.shape[0]):
                mask = gl[m] > 0
                assert np.all(a[m][~mask] == 0), f"{name}/{sampler}: material {m} leaked into pixels where the reference has none"
                if mask.any():
                    err = cases.rel_err(a[m], gl[m])[mask].max()
                    assert err <= LINE_RTOL, f"{name}/{sampler} view {i} material {m}: line integral rel err {err:.2e}"
            err = cases.rel_err(im, gi).max()
            assert err <= INT_RTOL, f"{name}/{sampler} view {i}: intensity rel err {err:.2e}"


@pytest.mark.parametrize("name", ["c1", "thorax_small"])
def test_single_volume_vs_reference_kernel_goldens(name):
    _compare_with_golden(name, ("alu", "tex", "hybrid"))


@pytest.mark.parametrize("name", ["multivol3", "multivol2_sameprio"])
def test_multi_volume_vs_reference_kernel_goldens(name):
    # priorities, same-priority averaging and the shared label cache quirk, every sampler
    _compare_with_golden(name, ("alu", "tex", "hybrid"))


def test_full_size_c2_vs_reference_kernel_golden():
    """BASELINE config 2: 512x512x400 CT -> 1536^2 detector, view 0, every 8th pixel stored."""
    _compare_with_golden("c2", ("hybrid", "alu"), views=[0])


def test_without_coefficient_records_everything_runs_on_the_texture_unit():
    """What the library does by itself when the 32 B / voxel records of the FMA-pipe sampler do not fit in device memory:
    the scene is sampled by the texture unit alone, single- and multi-volume, and still matches the reference goldens."""
    for name in ("thorax_small", "multivol3"):
        volumes, spectrum, priorities = cases.scene(name)
        g = cases.golden(name)
        W, H, sub = int(g["W"]), int(g["H"]), int(g["sub"])
        with Projector(volumes, priorities=priorities, spectrum=spectrum, step=float(g["step"]), neglog=False,
                       camera_intrinsics=geo.CameraIntrinsicTransform.from_sizes((W, H), 1.0, 1000.0), sampler="hybrid",
                       coefficient_records=False) as p:
            area = p.project_arrays(g["w2i_0"].reshape(1, 9), g["src_0"].reshape(1, -1, 3), g["ijk_0"].reshape(1, -1, 12), (W, H),
                                    float(g["max_ray_length"]), want="area")[0]
        gl = g["lineint_0"]
        a = area[:, ::sub, ::sub]
        for m in range(a.shape[0]):
            mask = gl[m] > 0
            if mask.any():
                assert cases.rel_err(a[m], gl[m])[mask].max() <= LINE_RTOL, (name, m)


def test_per_ray_kernel_variant_matches_too():
    volumes, spectrum, priorities = cases.scene("c1")
    g = cases.golden("c1")
    W, H, sub = int(g["W"]), int(g["H"]), int(g["sub"])
    with _projector(volumes, spectrum, priorities, "hybrid", W, H) as p:
        p.set_kernel_variant(1)
        area = p.project_arrays(g["w2i_0"][None], g["src_0"][None], g["ijk_0"][None], (W, H), float(g["max_ray_length"]), want="area")[0]
    for m in range(area.shape[0]):
        mask = g["lineint_0"][m] > 0
        assert cases.rel_err(area[m][::sub, ::sub], g["lineint_0"][m])[mask].max() <= LINE_RTOL


def test_fresh_poses_vs_cpu_oracle_including_odd_sensor_sizes():
    """Poses that are not in any golden file; sensor sizes that are not multiples of the 8x4 warp tile."""
    vol_ = phantoms.thorax_volume((64, 64, 50), (6.4, 6.4, 8.0), seed=7)
    st = cases.tables([vol_], "60KV_AL35", None)
    for (W, H), seed in (((37, 23), 11), ((64, 48), 12), ((1, 1), 13)):
        carm = phantoms.MobileCArmGeometry(sensor_width=W, sensor_height=H, pixel_size=298.0 / max(W, H))
        poses = phantoms.c2_poses(2, seed=seed, carm=carm)
        with Projector(vol_, spectrum="60KV_AL35", step=0.25, neglog=False, camera_intrinsics=carm.camera_intrinsics, sampler="hybrid") as p:
            area = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
            img = p.project(*poses, max_ray_length=carm.max_ray_length)
        for n, pose in enumerate(poses):
            w2i, src, ijk = geo.pose_arrays(pose, [vol_])
            r = cpu_oracle.project([vol_.data], st.labels, st.M, W, H, 0.25, w2i, src, ijk, carm.max_ray_length, st.energies, st.pdf, st.mu)
            for m in range(st.M):
                mask = r.area[m] > 0
                if mask.any():
                    assert cases.rel_err(area[n, m], r.area[m])[mask].max() <= LINE_RTOL
                assert np.all(area[n, m][~mask] == 0)
            assert cases.rel_err(img[n], r.intensity).max() <= INT_RTOL


def test_live_reference_kernel_on_new_poses():
    """Runs the reference's own cubin (oracle/_ref, built from /root/reference by oracle/Makefile) next to
    the CUDA path on poses generated here."""
    from oracle import ref_gpu

    if not ref_gpu.available():
        pytest.skip("oracle/_ref not shipped")
    volumes, spectrum, priorities = cases.scene("thorax_small")
    st = cases.tables(volumes, spectrum, priorities)
    carm = phantoms.MobileCArmGeometry(sensor_width=200, sensor_height=160, pixel_size=1.49)
    poses = phantoms.c2_poses(3, seed=99, carm=carm)
    ref = ref_gpu.RefProjector([v.data for v in volumes], st.labels, st.M, lineint=True)
    refp = ref_gpu.RefProjector([v.data for v in volumes], st.labels, st.M)
    refp.set_spectrum(st.energies, st.pdf, st.mu)
    with Projector(volumes, spectrum=spectrum, neglog=False, camera_intrinsics=carm.camera_intrinsics, sampler="hybrid") as p:
        area = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
        img = p.project(*poses, max_ray_length=carm.max_ray_length)
    for n, pose in enumerate(poses):
        w2i, src, ijk = geo.pose_arrays(pose, volumes)
        li = ref.line_integrals(200, 160, 0.1, w2i, src, ijk, carm.max_ray_length)
        ri, _, _ = refp.project(200, 160, 0.1, w2i, src, ijk, carm.max_ray_length)
        for m in range(st.M):
            mask = li[m] > 0
            assert cases.rel_err(area[n, m], li[m])[mask].max() <= LINE_RTOL
        assert cases.rel_err(img[n], ri).max() <= INT_RTOL
    ref.close(); refp.close()


@pytest.mark.parametrize("priorities,enabled", [(None, (1, 1, 1)), ([0, 1, 2], (1, 1, 1)), ([2, 0, 1], (1, 0, 1))])
def test_live_reference_kernel_ct_plus_two_tool_volumes(priorities, enabled):
    """BASELINE config 3 in small: CT + two crossing K-wire volumes, every sampler, against the reference cubin run on the
    same poses.  Tiles that see one volume take the lock-step kernel, tiles over the wires its per-sample priority pick,
    and tiles the shared-label-cache test flags are replayed by the general kernel -- all three must agree with the
    reference, whatever the priorities, and with a volume switched off."""
    from oracle import ref_gpu

    if not ref_gpu.available():
        pytest.skip("oracle/_ref not shipped")
    volumes = phantoms.c3_scene((128, 128, 100), (3.2, 3.2, 4.0))
    for v, e in zip(volumes, enabled):
        v.enabled = bool(e)
    poses, sdd = phantoms.cone_poses(2, seed=5, sensor=160, pixel=0.6)
    k = poses[0].intrinsic
    st = cases.tables(volumes, "90KV_AL40", priorities)
    ref = ref_gpu.RefProjector([v.data for v in volumes], st.labels, st.M, lineint=True)
    pr = priorities if priorities is not None else [2, 1, 0]
    results = {}
    for sampler in ("alu", "tex", "hybrid"):
        with Projector(volumes, priorities=priorities, spectrum="90KV_AL40", neglog=False, camera_intrinsics=k,
                       source_to_detector_distance=sdd, sampler=sampler) as p:
            results[sampler] = p.project_line_integrals(*poses)
            mrl = p.max_ray_length
    for n, pose in enumerate(poses):
        w2i, src, ijk = geo.pose_arrays(pose, volumes)
        li = ref.line_integrals(160, 160, 0.1, w2i, src, ijk, mrl, priority=pr, enabled=list(enabled))
        iron = st.all_materials.index("iron")
        if any(enabled[w] and pr[w] < pr[0] for w in (1, 2)):
            assert (li[iron] > 0).sum() > 100, "a wire that outranks the CT must show up"
        else:
            assert not (li[iron] > 0).any()
        for sampler, area in results.items():
            for m in range(st.M):
                mask = li[m] > 0
                if mask.any():
                    assert cases.rel_err(area[n, m], li[m])[mask].max() <= LINE_RTOL, (sampler, n, m)
                assert np.all(area[n, m][~mask] == 0), (sampler, n, m)
    ref.close()


def test_live_reference_kernel_label_cache_coincidences():
    """SURVEY.md App. A Q3 at full strength: two volumes on the same voxel grid, offset by less than a voxel, so that the
    second volume's floored sample coordinates equal the first one's at most steps and the reference serves it the FIRST
    volume's labels.  The lock-step kernel must detect this per ray and hand those tiles to the step-by-step replay."""
    from oracle import ref_gpu

    if not ref_gpu.available():
        pytest.skip("oracle/_ref not shipped")
    a = phantoms.thorax_volume((64, 64, 50), (6.4, 6.4, 8.0), seed=7)
    b = phantoms.thorax_volume((64, 64, 50), (6.4, 6.4, 8.0), seed=8)
    b.translate((1.9, 1.2, 2.6))                      # 0.3 voxel in every axis
    volumes, priorities = [a, b], [1, 0]              # b outranks a wherever both are hit
    st = cases.tables(volumes, "60KV_AL35", priorities)
    carm = phantoms.MobileCArmGeometry(sensor_width=96, sensor_height=80, pixel_size=3.0)
    poses = phantoms.c2_poses(2, seed=11, carm=carm)
    ref = ref_gpu.RefProjector([v.data for v in volumes], st.labels, st.M, lineint=True)
    with Projector(volumes, priorities=priorities, spectrum="60KV_AL35", neglog=False, camera_intrinsics=carm.camera_intrinsics,
                   sampler="hybrid") as p:
        area = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
        p.set_kernel_variant(1)                        # the general kernel alone
        area_general = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
    # the scene is sensitive to the quirk: listing the same two volumes in the other order (b first: it then always
    # fetches its own labels) changes the picture
    with Projector([b, a], priorities=[0, 1], spectrum="60KV_AL35", neglog=False, camera_intrinsics=carm.camera_intrinsics) as p:
        swapped = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
    assert cases.rel_err(swapped, area)[area > 1.0].max() > 1e-3
    # a projector that ignored the quirk (own labels always) would be far off: b's bone sampled with a's labels etc.
    for n, pose in enumerate(poses):
        w2i, src, ijk = geo.pose_arrays(pose, volumes)
        li = ref.line_integrals(96, 80, 0.1, w2i, src, ijk, carm.max_ray_length, priority=priorities)
        for m in range(st.M):
            mask = li[m] > 0
            assert cases.rel_err(area[n, m], li[m])[mask].max() <= LINE_RTOL, (n, m)
            assert cases.rel_err(area_general[n, m], li[m])[mask].max() <= LINE_RTOL, (n, m)
    ref.close()


def test_lock_step_path_equals_step_by_step_replay_on_mixed_scenes():
    """Scenes the reference cubins of oracle/_ref do not cover (four volumes; volumes plus additive and subtractive
    meshes): the split lock-step / work-list path must reproduce the general kernel, which replays projectKernel step by
    step and is itself pinned against the reference on the smaller scenes above."""
    from deepdrr_b200.vol import Mesh

    ct = phantoms.thorax_volume((96, 96, 80), (4.2, 4.2, 5.0), seed=2)
    wires = []
    for i, (tip, axis) in enumerate([((-30.0, -40.0, 5.0), (0.3, 1.0, 0.1)), ((20.0, -50.0, -10.0), (-0.2, 1.0, 0.0)),
                                     ((0.0, -30.0, 25.0), (0.0, 1.0, -0.3))]):
        w = phantoms.kwire_volume(length_mm=70.0, spacing=0.3, half_width=5)
        phantoms.place_kwire(w, tip, axis)
        wires.append(w)
    sv, sf = phantoms.screw_mesh(rings_per_mm=0.5, segments=16)
    screw = Mesh(sv, sf, material="titanium")
    phantoms.place_kwire(screw, (10.0, -45.0, 0.0), (0.1, 1.0, 0.2))
    bv, bf = phantoms.icosphere(28.0, 2)
    ball = Mesh(bv, bf, material="lung", density=0.3, subtractive=True, layer=1)
    ball.translate((-25.0, 5.0, 10.0))
    poses, sdd = phantoms.cone_poses(3, seed=9, sensor=144, pixel=0.8)
    k = poses[0].intrinsic
    scenes = {"four volumes": ([ct] + wires, None), "four volumes, explicit priorities": ([ct] + wires, [3, 0, 2, 1]),
              "two volumes + meshes": ([ct, wires[0], screw, ball], None), "one volume + meshes": ([ct, screw, ball], None)}
    for name, (objs, pr) in scenes.items():
        res = {}
        for variant in (0, 1):
            with Projector(objs, priorities=pr, spectrum="90KV_AL40", neglog=False, camera_intrinsics=k, source_to_detector_distance=sdd) as p:
                p.set_kernel_variant(variant)
                res[variant] = p.project_line_integrals(*poses)
        a, b = res[0], res[1]
        assert a.shape == b.shape and np.isfinite(a).all()
        mask = b > 0
        assert np.all(a[~mask] == 0), name
        # both paths follow the reference's order of operations; they differ only where the texture-unit emulation of the
        # replay kernel and the texture unit itself disagree by an ulp
        assert cases.rel_err(a, b)[mask].max() <= LINE_RTOL, name
        assert (a == b).mean() > 0.5, name


# ---------------------------------------------------------------------------------------------
# properties and edge cases
# ---------------------------------------------------------------------------------------------
def test_density_scaling_is_exactly_linear_and_batch_equals_single():
    v = phantoms.thorax_volume((48, 48, 40), (8.5, 8.5, 10.0), seed=2)
    carm = phantoms.MobileCArmGeometry(sensor_width=96, sensor_height=64, pixel_size=3.0)
    poses = phantoms.c2_poses(3, seed=5, carm=carm)
    with Projector(v, spectrum="90KV_AL40", neglog=False, camera_intrinsics=carm.camera_intrinsics) as p:
        a1 = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
        singles = np.stack([p.project_line_integrals(q, max_ray_length=carm.max_ray_length)[0] for q in poses])
    assert np.array_equal(a1, singles)                       # a batch is exactly the per-view results
    v2 = phantoms.thorax_volume((48, 48, 40), (8.5, 8.5, 10.0), seed=2)
    v2.data *= 2.0                                            # power of two: every product stays exact
    with Projector(v2, spectrum="90KV_AL40", neglog=False, camera_intrinsics=carm.camera_intrinsics) as p:
        a2 = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
    assert np.array_equal(a2, 2.0 * a1)


def test_homogeneous_box_line_integral_is_rho_times_chord():
    """Analytic KAT: constant density -> line integral = rho * step * (#in-range samples - 1/2 per
    half-weighted end); here checked against rho * chord length to within one step."""
    n = 40
    hu = np.full((n, n, n), 40.0, dtype=np.float32)
    from deepdrr_b200 import Volume

    c = (n - 1) / 2.0
    v = Volume.from_hu(hu, anatomical_from_IJK=geo.FrameTransform(np.array([[1, 0, 0, -c], [0, 1, 0, -c], [0, 0, 1, -c], [0, 0, 0, 1.0]])))
    rho = float(v.data[0, 0, 0])
    proj, mrl = phantoms.c1_camera(16, direction=(0.0, 1.0, 0.0))
    with Projector(v, spectrum="90KV_AL40", neglog=False, camera_intrinsics=proj.intrinsic) as p:
        a = p.project_line_integrals(proj, max_ray_length=mrl)[0]
    mats = p.all_materials
    soft = a[mats.index("soft tissue")]
    # central ray: chord = n voxels = 40 mm -> rho * 4.0 g/cm^2, edge effects below one step (0.1 mm)
    assert abs(soft[8, 8] - rho * n / 10.0) <= rho * 0.011
    assert np.all(a[mats.index("bone")] == 0) and np.all(a[mats.index("air")] == 0)


def test_disabled_volume_and_missed_volume_give_unattenuated_beam():
    v = phantoms.c1_volume(32)
    proj, mrl = phantoms.c1_camera(24)
    st = cases.tables([v], "90KV_AL40", None)
    i0 = np.float32(0)
    pp0 = np.float32(0)
    for b in range(st.n_bins):  # project_kernel.cu:637-646 with zero area densities
        pp0 = np.float32(pp0 + st.pdf[b])
        i0 = np.float32(np.float32(st.energies[b]) * st.pdf[b] + i0)
    with Projector(v, spectrum="90KV_AL40", neglog=False, camera_intrinsics=proj.intrinsic) as p:
        v.enabled = False
        img = p.project(proj, max_ray_length=mrl)
        v.enabled = True
        away, _ = phantoms.c1_camera(24, direction=(0.0, 1.0, 0.0))
        away = phantoms.look_at_projection((0, 2000.0, 0), (0, 1.0, 0), (0, 0, 1), proj.intrinsic)  # looking away from the volume
        img2 = p.project(away, max_ray_length=mrl)
        p.neglog = True
        z = p.project(away, max_ray_length=mrl)
    assert np.allclose(img, i0, rtol=2e-6) and np.allclose(img2, i0, rtol=2e-6)
    assert np.all(z == 0)  # constant image -> neglog maps it to 0 (utils/image_utils.py:42-49)


def test_neglog_clip_and_batch_semantics_match_host_restatement():
    volumes, spectrum, priorities = cases.scene("c1")
    proj, mrl = phantoms.c1_camera(64)
    proj2, _ = phantoms.c1_camera(64, direction=(1.0, 0.2, 0.0))
    with Projector(volumes, spectrum=spectrum, neglog=False, camera_intrinsics=proj.intrinsic) as p:
        raw = p.project(proj, proj2, max_ray_length=mrl)
        p.neglog = True
        nl = p.project(proj, proj2, max_ray_length=mrl)
        p.intensity_upper_bound = float(np.median(raw))
        nlc = p.project(proj, proj2, max_ray_length=mrl)
    assert raw.shape == (2, 64, 64) and raw.dtype == np.float32
    assert np.allclose(nl, cpu_oracle.neglog(raw), atol=3e-6, rtol=0)
    assert nl.min() == 0.0 and nl.max() == 1.0
    assert np.allclose(nlc, cpu_oracle.neglog(np.minimum(raw, np.float32(np.median(raw)))), atol=3e-6, rtol=0)


def test_poisson_noise_statistics():
    """add_noise (analytic_generators.py:10-18) is unseeded NumPy in the reference -> statistical parity only:
    noise is zero-mean with variance sum(k^2) * I^2 / (photon_prob * photon_count) (3x3 blur kernel k)."""
    v = phantoms.c1_volume(32)
    proj, mrl = phantoms.c1_camera(128)
    arrays = [a[None] for a in geo.pose_arrays(proj, [v])]
    N = 2000
    with Projector(v, spectrum="90KV_AL40", neglog=False, camera_intrinsics=proj.intrinsic, add_noise=True, photon_count=N,
                   noise_seed=1234) as p:
        noisy = np.stack([p.project(proj, max_ray_length=mrl) for _ in range(32)])
        clean, pp = p.project_arrays(*arrays, (128, 128), mrl, want="intensity+photon_prob")
    clean, pp = clean[0].astype(np.float64), pp[0].astype(np.float64)
    assert np.all(noisy >= 0)
    d = (noisy - clean)[:, 8:-8, 8:-8]
    k2 = 0.03**2 + 0.06**2 + 0.02**2 + 0.11**2 + 0.98**2 + 0.11**2 + 0.02**2 + 0.06**2 + 0.03**2
    expect_var = (k2 * clean**2 / (pp * N))[8:-8, 8:-8]
    ratio = d.var(axis=0, ddof=1) / expect_var
    assert abs(ratio.mean() - 1.0) < 0.05, ratio.mean()
    assert abs((d / np.sqrt(expect_var)).mean()) < 0.01          # zero mean (in units of sigma)
    with Projector(v, spectrum="90KV_AL40", neglog=False, camera_intrinsics=proj.intrinsic, add_noise=True, photon_count=N,
                   noise_seed=1234) as q:
        again = q.project(proj, max_ray_length=mrl)
    assert np.array_equal(again, noisy[0])                       # same seed reproduces


def test_collected_energy_matches_formula():
    v = phantoms.c1_volume(32)
    carm = phantoms.MobileCArmGeometry(sensor_width=48, sensor_height=40, pixel_size=6.0)
    pose = carm.camera_projection(0.2, -0.1, (0, 0, 0))

    class Dev:
        source_to_detector_distance = carm.source_to_detector_distance
        camera_intrinsics = carm.camera_intrinsics
        detector_height, detector_width = carm.detector_height, carm.detector_width

        def get_camera_projection(self):
            return pose

    w2i, src, ijk = geo.pose_arrays(pose, [v])
    with Projector(v, device=Dev(), neglog=False, collected_energy=True, photon_count=5000) as p:
        ce = p.project()
        p.collected_energy = False
        inten = p.project()
    r = cpu_oracle.project([v.data], cases.tables([v], "90KV_AL40", None).labels, 3, 48, 40, 0.1, w2i, src, ijk, carm.max_ray_length,
                           *[getattr(cases.tables([v], "90KV_AL40", None), k) for k in ("energies", "pdf", "mu")], want_solid=True)
    px = carm.source_to_detector_distance / carm.camera_intrinsics.fx
    expect = inten.astype(np.float64) * r.solid * 5000 / r.solid.astype(np.float64).mean() / (px * px)   # projector.py:846-852
    assert np.allclose(ce, expect, rtol=2e-5)


def test_lifecycle_errors_and_reuse():
    v = phantoms.c1_volume(16)
    proj, mrl = phantoms.c1_camera(16)
    p = Projector(v, camera_intrinsics=proj.intrinsic)
    with pytest.raises(RuntimeError):
        p.project(proj)
    p.initialize()
    with pytest.raises(RuntimeError):
        p.initialize()                                        # projector.py:1397-1398
    with pytest.raises(ValueError):
        p.project()                                           # no pose and no device (projector.py:632-635)
    a = p.project(proj, max_ray_length=mrl)
    assert a.shape == (16, 16)                                # one view -> [H, W] (projector.py:704-705)
    p.free()
    p.free()                                                  # idempotent
    for _ in range(10):                                       # re-create / leak check (tests/test_core.py:896-898)
        with Projector(v, camera_intrinsics=proj.intrinsic) as q:
            b = q.project(proj, max_ray_length=mrl)
        assert np.array_equal(a, b)


def test_outside_air_attenuation_vs_live_reference_kernel():
    """attenuate_outside_volume=True (project_kernel.cu:359-361, 522-527, 558-560), one and two volumes."""
    from oracle import ref_gpu

    if not ref_gpu.available():
        pytest.skip("oracle/_ref not shipped")
    vs = phantoms.thorax_volume((64, 64, 50), (6.4, 6.4, 8.0), seed=7)
    vs2 = phantoms.thorax_volume((48, 48, 40), (6.4, 6.4, 8.0), seed=9)
    vs2.translate((40.0, -30.0, 25.0))
    carm = phantoms.MobileCArmGeometry(sensor_width=96, sensor_height=72, pixel_size=3.1)
    poses = phantoms.c2_poses(2, seed=21, carm=carm)
    for volumes in ([vs], [vs, vs2]):
        with Projector(volumes, spectrum="90KV_AL40", neglog=False, camera_intrinsics=carm.camera_intrinsics, attenuate_outside_volume=True) as p:
            assert p.all_materials == ["air", "bone", "soft tissue"] and p.air_index == 0
            area = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
            labels = [__import__("deepdrr_b200.scene", fromlist=["remap_labels"]).remap_labels(v, p.all_materials) for v in volumes]
        ref = ref_gpu.RefProjector([v.data for v in volumes], labels, 3, lineint=True, variant="att0")
        for n, pose in enumerate(poses):
            w2i, src, ijk = geo.pose_arrays(pose, volumes)
            li = ref.line_integrals(96, 72, 0.1, w2i, src, ijk, carm.max_ray_length)
            # the outside-air terms are active everywhere -- and negative: (ray_length - maxAlpha) / step with the
            # ray_length ~ 1 of the R^T K^-1 convention; reproduced as is
            assert (li[0] != 0).all()
            for m in range(3):
                mask = li[m] != 0
                assert np.all(area[n, m][~mask] == 0)
                if mask.any():
                    assert cases.rel_err(area[n, m], li[m])[mask].max() <= LINE_RTOL
        ref.close()


def test_gpu_side_hu_preparation_is_bit_identical_to_the_host_path():
    """drr_add_volume_hu (HU -> density + threshold labels on the device) vs Volume.from_hu on the host."""
    from deepdrr_b200 import HUVolume, Volume

    hu = phantoms.thorax_hu((40, 48, 36), (10.0, 8.5, 11.0), seed=5)
    hu[3, 4, 5] = np.nan                                                            # unlabeled voxel -> id 0 (air), density NaN-propagating max
    hu[0, 0, 0], hu[1, 1, 1], hu[2, 2, 2] = -800.0, 350.0, 350.00003               # threshold edges
    hu[3, 4, 5] = -800.00006
    host = Volume.from_hu(hu, anatomical_from_IJK=geo.FrameTransform.from_scaling(8.0, (-160, -190, -150)))
    dev_ = HUVolume(hu, anatomical_from_IJK=geo.FrameTransform.from_scaling(8.0, (-160, -190, -150)))
    carm = phantoms.MobileCArmGeometry(sensor_width=64, sensor_height=48, pixel_size=4.5)
    poses = phantoms.c2_poses(2, seed=31, carm=carm)
    out = []
    for v in (host, dev_):
        with Projector(v, spectrum="90KV_AL40", neglog=False, camera_intrinsics=carm.camera_intrinsics, step=0.5) as p:
            out.append(p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length))
    assert dev_._host is None
    assert np.array_equal(out[0], out[1])


def test_pipelined_host_batches_equal_single_piece_batches():
    """drr_project projects a host-bound batch of >= 4 views in two halves so that the copy of the first runs under the march of
    the second.  Same bits as in one piece -- including utils.neglog's quirk that ONE constant image zeroes the WHOLE batch
    (utils/image_utils.py:42-49), whichever half it is in."""
    vol_ = phantoms.thorax_volume((64, 64, 50), (6.4, 6.4, 8.0), seed=7)
    carm = phantoms.MobileCArmGeometry(sensor_width=96, sensor_height=80, pixel_size=3.1)
    poses = phantoms.c2_poses(7, seed=31, carm=carm)
    away = phantoms.look_at_projection((0.0, 0.0, 5000.0), (0.0, 0.0, 1.0), (0, 1, 0), carm.camera_intrinsics)   # sees nothing: constant image
    with Projector(vol_, spectrum="120KV_AL43", neglog=True, camera_intrinsics=carm.camera_intrinsics) as p:
        for batch in (poses, poses[:4], poses[:3] + [away] + poses[3:6], [away] + poses[:5], poses[:5] + [away]):
            p.set_pipeline(True)
            a = p.project(*batch, max_ray_length=carm.max_ray_length).copy()
            p.set_pipeline(False)
            b = p.project(*batch, max_ray_length=carm.max_ray_length).copy()
            assert np.array_equal(a, b)
            if any(q is away for q in batch):
                assert not a.any()
            else:
                assert a.min() == 0.0 and a.max() == 1.0
        p.neglog = False
        p.set_pipeline(True)
        raw = p.project(*poses, max_ray_length=carm.max_ray_length).copy()
        p.set_pipeline(False)
        assert np.array_equal(raw, p.project(*poses, max_ray_length=carm.max_ray_length))


def test_lane_layout_does_not_change_the_images():
    """DRR_TUNE_LANE_QUADS only changes which lane of a warp walks which pixel of its 8 x 8 tile (2 x 2 groups for the texture unit, or
    4 x 1 runs): every pixel's ray, its samples and their order stay the same, so the images must be identical -- on odd detector
    sizes (partial tiles) as well, and under all three samplers."""
    vol_ = phantoms.thorax_volume((64, 64, 50), (6.4, 6.4, 8.0), seed=7)
    for (w, h) in ((96, 80), (101, 67)):
        carm = phantoms.MobileCArmGeometry(sensor_width=w, sensor_height=h, pixel_size=3.1 * 96 / w)
        poses = phantoms.c2_poses(3, seed=5, carm=carm)
        for sampler in ("hybrid", "tex", "alu"):
            with Projector(vol_, spectrum="120KV_AL43", neglog=False, camera_intrinsics=carm.camera_intrinsics, sampler=sampler) as p:
                out = []
                for mode in (0, 1, 2):
                    p.set_lane_quads(mode)
                    out.append(p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length).copy())
                assert out[0].any()
                assert np.array_equal(out[0], out[1]) and np.array_equal(out[1], out[2]), (w, h, sampler)
                # The number of rays a lane walks through each staged box (DRR_TUNE_RAYS_PER_LANE: the library picks 1 or 2 from the
                # ray spacing) changes the boxes and with them where the segments begin, i.e. which steps of a ray the hybrid
                # sampler gives to the texture unit and which to its FMA-pipe emulation (equal in 99.8 % of fetches, 1 ulp apart
                # otherwise): the samples and their order stay, so the texture sampler must agree exactly and the others to a few
                # 1e-7 -- far inside what either has against the reference.
                ref = {}
                for rays in (1, 2, 0):
                    p.set_rays_per_lane(rays)
                    ref[rays] = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length).copy()
                if sampler == "tex":
                    assert np.array_equal(ref[1], ref[2])
                ok = ref[2] > 0
                assert (ok == (ref[1] > 0)).all()
                assert (np.abs(ref[1] - ref[2])[ok] / ref[2][ok]).max() <= 2e-6, (w, h, sampler)
                assert np.array_equal(ref[0], ref[1]) or np.array_equal(ref[0], ref[2])


def test_fma_pipe_sampler_on_boxes_that_start_in_the_clamped_layers():
    """Regression: scenes 110, 170, 815 and 1115 of `tools/fuzz_single.py 1500 11`.  The general-segment path of the FMA-pipe sampler
    took a sample's fraction as (x - box origin) - floor(...), which is not exact when the staged box starts in the clamped cell
    layers (origin <= 0): a fraction of 191.49994 / 256 became 191.5000013 / 256, one fixed-point step of the texture unit's
    coordinate, 0.9 % of that sample -- 3.5e-5 of a pure-air pixel's line integral.  Now x - floor(x), exact."""
    from oracle import ref_gpu

    if not ref_gpu.available():
        pytest.skip("oracle/_ref not shipped")
    rng = np.random.default_rng(11)
    wanted = {110, 170, 815, 1115}
    for it in range(1116):
        sc = cases.random_single_volume_scene(rng, it, build=it in wanted)
        if it not in wanted:
            continue
        v, st, W, H = sc["volume"], sc["tables"], sc["W"], sc["H"]
        w2i, src, ijk = geo.pose_arrays(sc["pose"], [v])
        ref = ref_gpu.RefProjector([v.data], st.labels, st.M, lineint=True)
        li = ref.line_integrals(W, H, 0.1, w2i, src, ijk, sc["mrl"])
        ref.close()
        for sampler in ("alu", "hybrid"):
            with Projector(v, spectrum="90KV_AL40", neglog=False, camera_intrinsics=sc["k"], source_to_detector_distance=sc["sdd"], sampler=sampler) as p:
                area = p.project_line_integrals(sc["pose"], max_ray_length=sc["mrl"])[0]
            for m in range(st.M):
                mask = li[m] > 0
                assert np.all(area[m][~mask] == 0)
                if mask.any():
                    err = cases.rel_err(area[m], li[m])[mask].max()
                    assert err <= 2e-6, f"scene {it} [{sampler}] material {m}: {err:.2e}"
