"""``calculate_solid_angle`` (reference project_kernel.cu:14-133, reached through a non-null ``solid_angle`` pointer at :213-216)
pinned to the reference kernel itself: tests/golden/solid_angle.npz holds what the unmodified cubin wrote for five cameras
(generator: tools/make_goldens.py, ``solid_angle_case``).  The CPU oracle's restatement is checked here without a GPU; the CUDA
path (``solid_angle_kernel`` inside ``collected_energy``) is checked against the same vectors and, live, against the cubin."""
import os

import numpy as np
import pytest

import cases
from deepdrr_b200 import Projector, geo, phantoms
from oracle import cpu_oracle

GOLD = os.path.join(cases.GOLDEN, "solid_angle.npz")


def _cameras():
    cams = [phantoms.c1_camera(direction=d) for d in ((0.3, 1.0, 0.2), (0.0, 1.0, 0.0), (1.0, 0.0, 1.0))]
    carm = phantoms.MobileCArmGeometry(sensor_width=192, sensor_height=160, pixel_size=1.5)
    cams.append((carm.camera_projection(0.5, -0.4, (10.0, -20.0, 5.0)), carm.max_ray_length))
    k = geo.CameraIntrinsicTransform(np.array([[900.0, 0, 20.0], [0, 1100.0, 150.0], [0, 0, 1]]), sensor_height=120, sensor_width=100)
    cams.append((phantoms.look_at_projection((100.0, -700.0, 50.0), (-0.1, 1.0, 0.0), (0, 0, 1), k), 3000.0))
    return cams


def test_oracle_solid_angle_matches_reference_kernel():
    g = np.load(GOLD)
    v = phantoms.c1_volume(16)
    st = cases.tables([v], "90KV_AL40", None)
    for i, (proj, mrl) in enumerate(_cameras()):
        W, H = int(g[f"W_{i}"]), int(g[f"H_{i}"])
        w2i, src, ijk = geo.pose_arrays(proj, [v])
        assert np.array_equal(w2i, g[f"w2i_{i}"]), "the camera recipe changed: regenerate tests/golden/solid_angle.npz"
        r = cpu_oracle.project([v.data], st.labels, st.M, W, H, 1.0e6, w2i, src, ijk, mrl, st.energies, st.pdf, st.mu, want_solid=True)
        want = g[f"solid_{i}"]
        assert want.shape == (H, W) and np.all(want > 0)
        # glibc atan2f against CUDA atan2f: a couple of ulps
        assert cases.rel_err(r.solid, want).max() <= 2e-6, (i, cases.rel_err(r.solid, want).max())
        # sanity of the pinned vectors themselves: the pixels tile the detector's solid angle, largest on the optical axis
        assert abs(float(want.sum(dtype=np.float64)) - float(r.solid.sum(dtype=np.float64))) <= 1e-6 * float(want.sum(dtype=np.float64))


@pytest.mark.gpu
def test_cuda_collected_energy_uses_the_reference_solid_angle():
    """projector.py:833-853: collected = intensity * solid_angle * photon_count / mean(solid_angle) / pixel_area.  The solid angle of
    the CUDA path is recovered from collected / intensity and compared with the reference kernel's own buffer."""
    g = np.load(GOLD)
    v = phantoms.c1_volume(16)
    for i, (proj, mrl) in enumerate(_cameras()[:4]):
        W, H = int(g[f"W_{i}"]), int(g[f"H_{i}"])
        with Projector(v, camera_intrinsics=proj.intrinsic, source_to_detector_distance=1000.0, neglog=False, collected_energy=True,
                       photon_count=7000) as p:
            ce = p.project(proj, max_ray_length=mrl)
            p.collected_energy = False
            inten = p.project(proj, max_ray_length=mrl)
        k = proj.intrinsic
        px = (1000.0 / k.fx) * (1000.0 / k.fy)
        want = g[f"solid_{i}"].astype(np.float64)
        expect = inten.astype(np.float64) * want * 7000 / want.mean() / px
        assert cases.rel_err(ce, expect).max() <= 2e-5, (i, cases.rel_err(ce, expect).max())


@pytest.mark.gpu
def test_live_reference_solid_angle_on_a_fresh_camera():
    from oracle import ref_gpu

    if not ref_gpu.available():
        pytest.skip("oracle/_ref not shipped")
    v = phantoms.c1_volume(16)
    st = cases.tables([v], "90KV_AL40", None)
    carm = phantoms.MobileCArmGeometry(sensor_width=140, sensor_height=100, pixel_size=2.0)
    proj = carm.camera_projection(-0.6, 0.3, (5.0, 0.0, -10.0))
    w2i, src, ijk = geo.pose_arrays(proj, [v])
    ref = ref_gpu.RefProjector([v.data], st.labels, st.M)
    ref.set_spectrum(st.energies, st.pdf, st.mu)
    want = ref.solid_angle(140, 100, w2i, src, ijk, carm.max_ray_length).astype(np.float64)
    ref.close()
    r = cpu_oracle.project([v.data], st.labels, st.M, 140, 100, 1.0e6, w2i, src, ijk, carm.max_ray_length, st.energies, st.pdf, st.mu, want_solid=True)
    assert cases.rel_err(r.solid, want).max() <= 2e-6
    with Projector(v, camera_intrinsics=carm.camera_intrinsics, source_to_detector_distance=carm.source_to_detector_distance, neglog=False,
                   collected_energy=True, photon_count=100) as p:
        ce = p.project(proj, max_ray_length=carm.max_ray_length)
        p.collected_energy = False
        inten = p.project(proj, max_ray_length=carm.max_ray_length)
    px = (carm.source_to_detector_distance / carm.camera_intrinsics.fx) ** 2
    assert cases.rel_err(ce, inten.astype(np.float64) * want * 100 / want.mean() / px).max() <= 2e-5
