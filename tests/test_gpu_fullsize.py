"""Full-size parity on what bench.py times, against the reference's own kernel run LIVE next to the CUDA path
(oracle/_ref, built from /root/reference by oracle/Makefile; the tests skip when it was not shipped):

* BASELINE config 2 at 512x512x400 -> 1536^2 on three of the bench's own poses (``c2_poses(1000, seed=1)`` numbers 1, 500
  and 999), EVERY pixel, samplers ``hybrid``, ``tex`` and ``coefficient_records=False``;
* BASELINE config 3 at full size: the 512x512x400 CT plus two 21x21x2000 K-wire volumes at 0.1 mm, 384^2 sensor;
* single fine-grid volumes seen from more than a metre away, where the drift of the accumulated fp32 ``alpha`` exceeds the
  fixed slack the lock-step kernel used to bound its staged boxes with (VERDICT r1 "lock-step slack hole").

Tolerances are north_star's: line integrals 1e-5 relative per pixel and material (no floor), intensity 1e-4 relative.
"""
import numpy as np
import pytest

import cases
from deepdrr_b200 import Projector, geo, phantoms

pytestmark = pytest.mark.gpu

LINE_RTOL = 1e-5
INT_RTOL = 1e-4


def _ref():
    from oracle import ref_gpu

    if not ref_gpu.available():
        pytest.skip("oracle/_ref not shipped")
    return ref_gpu


def _check(area, img, li, ri, tag):
    for m in range(li.shape[0]):
        mask = li[m] > 0
        assert np.all(area[m][~mask] == 0), f"{tag}: material {m} leaked into pixels where the reference has none"
        if mask.any():
            err = cases.rel_err(area[m], li[m])[mask].max()
            assert err <= LINE_RTOL, f"{tag} material {m}: line integral rel err {err:.2e}"
    if img is not None:
        err = cases.rel_err(img, ri).max()
        assert err <= INT_RTOL, f"{tag}: intensity rel err {err:.2e}"


def test_c2_bench_poses_every_pixel_vs_live_reference():
    ref_gpu = _ref()
    carm = phantoms.MobileCArmGeometry()
    W, H = carm.sensor_width, carm.sensor_height
    vol = phantoms.thorax_volume()
    st = cases.tables([vol], "120KV_AL43", None)
    all_poses = phantoms.c2_poses(1000, seed=1, carm=carm)          # bench.py's pose set
    ids = (1, 500, 999)
    poses = [all_poses[i] for i in ids]
    refl = ref_gpu.RefProjector([vol.data], st.labels, st.M, lineint=True)
    refp = ref_gpu.RefProjector([vol.data], st.labels, st.M)
    refp.set_spectrum(st.energies, st.pdf, st.mu)
    want = []
    for pose in poses:
        w2i, src, ijk = geo.pose_arrays(pose, [vol])
        li = refl.line_integrals(W, H, 0.1, w2i, src, ijk, carm.max_ray_length)
        ri, _, _ = refp.project(W, H, 0.1, w2i, src, ijk, carm.max_ray_length)
        want.append((li, ri))
    refl.close(); refp.close()
    for label, kw in (("hybrid", dict(sampler="hybrid")), ("tex", dict(sampler="tex")),
                      ("no records", dict(sampler="hybrid", coefficient_records=False))):
        with Projector(vol, spectrum="120KV_AL43", step=0.1, neglog=False, camera_intrinsics=carm.camera_intrinsics,
                       source_to_detector_distance=carm.source_to_detector_distance, **kw) as p:
            area = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
            img = p.project(*poses, max_ray_length=carm.max_ray_length)
            assert p.launch_count() > 0
        for n, i in enumerate(ids):
            _check(area[n], img[n], want[n][0], want[n][1], f"C2 pose {i} [{label}]")


def test_c3_full_size_vs_live_reference():
    ref_gpu = _ref()
    volumes = phantoms.c3_scene()                                    # 512x512x400 CT + two 21x21x2000 K-wires (0.1 mm)
    st = cases.tables(volumes, "120KV_AL43", None)
    poses, sdd = phantoms.cone_poses(2, seed=3)                      # 384^2 at 0.3 mm, SDD 1000 (README.md:78-83)
    k = poses[0].intrinsic
    refl = ref_gpu.RefProjector([v.data for v in volumes], st.labels, st.M, lineint=True)
    with Projector(volumes, spectrum="120KV_AL43", neglog=False, camera_intrinsics=k, source_to_detector_distance=sdd) as p:
        area = p.project_line_integrals(*poses)
        mrl = p.max_ray_length
    iron = st.all_materials.index("iron")
    for n, pose in enumerate(poses):
        w2i, src, ijk = geo.pose_arrays(pose, volumes)
        li = refl.line_integrals(384, 384, 0.1, w2i, src, ijk, mrl, priority=st.priorities)
        assert (li[iron] > 0).sum() > 500, "the wires must be in view"
        _check(area[n], None, li, None, f"C3 view {n}")
    refl.close()


@pytest.mark.parametrize("spacing,distance", [(0.1, 1100.0), (0.05, 1100.0), (0.02, 1500.0), (0.1, 4200.0)])
def test_fine_grid_far_source_single_volume_vs_live_reference(spacing, distance):
    """A K-wire volume on a 0.1 / 0.05 / 0.02 mm grid (reference: vol/kwire.py:90-102) seen from 1.1 - 4.2 m through a detector fine
    enough for the lock-step kernel to be picked: per step, alpha moves by ``step`` rounded to an ulp of 1.2e-4 .. 4.9e-4 mm, i.e.
    after a 32-step segment the sample can sit 0.02 .. 0.4 voxel from where ``alpha + 32 * step`` puts it -- more than the 0.01
    voxel the staged box used to allow.  The library now sizes that slack per launch (drr_capi.cu: march_slack) and hands scenes
    beyond a quarter voxel to the per-ray kernel."""
    ref_gpu = _ref()
    wire = phantoms.kwire_volume(length_mm=1000 * spacing, radius_mm=6 * spacing, tip_mm=30 * spacing, spacing=spacing, half_width=10)
    axis = np.array([0.05, 0.1, 1.0]) / np.linalg.norm([0.05, 0.1, 1.0])
    phantoms.place_kwire(wire, (0.0, 0.0, 0.0), axis)
    middle = 500 * spacing * axis                                   # a point on the wire's axis, half way along
    st = cases.tables([wire], "90KV_AL40", None)
    W, H = 96, 64
    sdd = distance + 20.0
    k = geo.CameraIntrinsicTransform.from_sizes((W, H), 4.0 * spacing * sdd / distance / 10.0, sdd)   # ~0.4 voxel between rays
    refl = ref_gpu.RefProjector([wire.data], st.labels, st.M, lineint=True)
    for direction in ((1.0, 0.2, 0.1), (0.3, 1.0, 0.45)):
        d = np.asarray(direction) / np.linalg.norm(direction)
        pose = phantoms.look_at_projection(middle - distance * d, d, (0, 0, 1), k)
        w2i, src, ijk = geo.pose_arrays(pose, [wire])
        li = refl.line_integrals(W, H, 0.1, w2i, src, ijk, distance + 500.0)
        assert (li[0] > 0).sum() > 200, "the wire must be in view"
        for sampler in ("hybrid", "alu", "tex"):
            with Projector(wire, spectrum="90KV_AL40", step=0.1, neglog=False, camera_intrinsics=k, source_to_detector_distance=sdd,
                           sampler=sampler) as p:
                area = p.project_line_integrals(pose, max_ray_length=distance + 500.0)
            _check(area[0], None, li, None, f"wire {spacing} mm from {distance} mm [{sampler}] dir {direction}")
    refl.close()


@pytest.mark.parametrize("step", [0.3, 1.0, 2.5, 7.0])
def test_long_steps_vs_live_reference(step):
    """Steps from a third of a voxel to several voxels (1 mm voxels).  The lock-step kernels stage the cells of up to 32 steps at a
    time and have to cut their segments down when the steps are long; from about a voxel per step the host hands the scene to the
    per-ray kernel (drr_capi.cu: pick_variant).  Either way the samples are the reference's."""
    ref_gpu = _ref()
    vol = phantoms.thorax_volume((160, 128, 120), (1.0, 1.0, 1.0), seed=3)
    st = cases.tables([vol], "90KV_AL40", None)
    W, H = 320, 256
    k = geo.CameraIntrinsicTransform.from_sizes((W, H), 0.6, 1000.0)   # rays 0.3 voxel apart at the volume
    refl = ref_gpu.RefProjector([vol.data], st.labels, st.M, lineint=True)
    poses = [phantoms.look_at_projection(-500.0 * np.asarray(d) / np.linalg.norm(d), np.asarray(d) / np.linalg.norm(d), (0, 0, 1), k)
             for d in ((0.3, 1.0, 0.2), (1.0, 0.1, -0.4))]
    for i, pose in enumerate(poses):
        w2i, src, ijk = geo.pose_arrays(pose, [vol])
        li = refl.line_integrals(W, H, step, w2i, src, ijk, 1200.0)
        assert (li > 0).any()
        for sampler in ("hybrid", "tex", "alu"):
            with Projector(vol, spectrum="90KV_AL40", step=step, neglog=False, camera_intrinsics=k, source_to_detector_distance=1000.0,
                           sampler=sampler) as p:
                for variant in (0, 1):   # 0: the library's choice, 1: the per-ray kernel
                    p.set_kernel_variant(variant)
                    area = p.project_line_integrals(pose, max_ray_length=1200.0)
                    _check(area[0], None, li, None, f"step {step} mm pose {i} [{sampler}, variant {variant}]")
    refl.close()


def test_five_volumes_vs_live_reference():
    """More volumes than the lock-step kernels are built for (4): the step-by-step kernel's wide instantiation
    (csrc/drr_march.cu, DRR_MAX_VOLUMES x DRR_MAX_MATERIALS register arrays, counts read at run time) against the reference
    compiled with -D NUM_VOLUMES=5 (projector.py:365-386 compiles for any count)."""
    ref_gpu = _ref()
    ct = phantoms.thorax_volume((96, 96, 80), (4.2, 4.2, 5.0), seed=2)
    wires = []
    for tip, axis in (((-30.0, -40.0, 5.0), (0.3, 1.0, 0.1)), ((20.0, -50.0, -10.0), (-0.2, 1.0, 0.0)), ((0.0, -30.0, 25.0), (0.0, 1.0, -0.3)),
                      ((-10.0, -45.0, -20.0), (0.5, 1.0, 0.4))):
        w = phantoms.kwire_volume(length_mm=70.0, spacing=0.3, half_width=5)
        phantoms.place_kwire(w, tip, axis)
        wires.append(w)
    volumes = [ct] + wires
    for priorities in (None, [4, 0, 3, 1, 2]):
        st = cases.tables(volumes, "90KV_AL40", priorities)
        poses, sdd = phantoms.cone_poses(2, seed=9, sensor=144, pixel=0.8)
        refl = ref_gpu.RefProjector([v.data for v in volumes], st.labels, st.M, lineint=True)
        with Projector(volumes, priorities=priorities, spectrum="90KV_AL40", neglog=False, camera_intrinsics=poses[0].intrinsic,
                       source_to_detector_distance=sdd) as p:
            area = p.project_line_integrals(*poses)
            mrl = p.max_ray_length
        for n, pose in enumerate(poses):
            w2i, src, ijk = geo.pose_arrays(pose, volumes)
            li = refl.line_integrals(144, 144, 0.1, w2i, src, ijk, mrl, priority=st.priorities)
            assert (li[st.all_materials.index("iron")] > 0).sum() > 100
            _check(area[n], None, li, None, f"five volumes, priorities {priorities}, view {n}")
        refl.close()


def test_nine_materials_vs_live_reference():
    """More materials than the templated kernels cover (8): one volume segmented into nine materials, reference compiled with
    -D NUM_MATERIALS=9."""
    from deepdrr_b200.vol import Volume

    ref_gpu = _ref()
    names = ["air", "blood", "bone", "concrete", "copper", "iron", "lung", "muscle", "soft tissue"]
    rng = np.random.default_rng(5)
    shape = (40, 36, 30)
    # blobs of materials: labels vary smoothly enough for uniform cells to exist, densities are noisy
    ii, jj, kk = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    labels = ((ii // 7 + 2 * (jj // 6) + 3 * (kk // 8)) % 9).astype(np.uint16)
    data = (0.5 + 0.2 * labels + rng.normal(0, 0.03, shape)).astype(np.float32).clip(0.0, None)
    a = np.diag([3.0, 3.0, 4.0, 1.0])
    a[:3, 3] = [-3.0 * (shape[0] - 1) / 2, -3.0 * (shape[1] - 1) / 2, -4.0 * (shape[2] - 1) / 2]
    vol = Volume(data, ({n: i for i, n in enumerate(names)}, labels), anatomical_from_IJK=geo.FrameTransform(a))
    st = cases.tables([vol], "90KV_AL40", None)
    assert st.M == 9
    carm = phantoms.MobileCArmGeometry(sensor_width=120, sensor_height=96, pixel_size=1.6)
    poses = phantoms.c2_poses(2, seed=21, carm=carm)
    refl = ref_gpu.RefProjector([vol.data], st.labels, st.M, lineint=True)
    refp = ref_gpu.RefProjector([vol.data], st.labels, st.M)
    refp.set_spectrum(st.energies, st.pdf, st.mu)
    with Projector(vol, spectrum="90KV_AL40", neglog=False, camera_intrinsics=carm.camera_intrinsics) as p:
        area = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
        img = p.project(*poses, max_ray_length=carm.max_ray_length)
    for n, pose in enumerate(poses):
        w2i, src, ijk = geo.pose_arrays(pose, [vol])
        li = refl.line_integrals(120, 96, 0.1, w2i, src, ijk, carm.max_ray_length)
        ri, _, _ = refp.project(120, 96, 0.1, w2i, src, ijk, carm.max_ray_length)
        assert sum(int((li[m] > 0).any()) for m in range(9)) == 9
        _check(area[n], img[n], li, ri, f"nine materials view {n}")
    refl.close(); refp.close()


def test_limits_are_the_header_limits():
    """include/drr_b200.h: DRR_MAX_VOLUMES 8, DRR_MAX_MATERIALS 16.  Beyond them the library says so when the volume / spectrum is
    handed over, not at the first projection."""
    from deepdrr_b200.vol import Volume
    from deepdrr_b200 import _lib

    tiny = np.ones((4, 4, 4), dtype=np.float32)
    vols = []
    for i in range(_lib.MAX_VOLUMES + 1):
        v = Volume(tiny, ({"bone": 0}, np.zeros(tiny.shape, np.uint16)))
        v.translate((6.0 * i, 0.0, 0.0))
        vols.append(v)
    k = geo.CameraIntrinsicTransform.from_sizes((16, 16), 1.0, 1000.0)
    with pytest.raises(ValueError):
        Projector(vols, camera_intrinsics=k).initialize()
    with Projector(vols[:8], camera_intrinsics=k, neglog=False) as p:            # eight volumes do run
        pose = phantoms.look_at_projection((20.0, -300.0, 0.0), (0, 1.0, 0), (0, 0, 1), k)
        a = p.project_line_integrals(pose, max_ray_length=2000.0)
        assert a.shape == (1, 1, 16, 16) and a.max() > 0
