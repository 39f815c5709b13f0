"""Monte Carlo scatter (BASELINE config 5).  The reference has no scatter implementation to compare with
(projector.py:530-531 raises; SURVEY.md App. C: parity unpinned), so the kernel is validated physically and
statistically, with the confidence bounds written next to each check:

  * energy bookkeeping closes (emitted = missed + absorbed + left unscattered + scattered),
  * the unscattered fraction through a homogeneous water slab equals exp(-L / mfp_total) within 4 sigma,
  * Rayleigh : Compton event counts follow the inverse mean free paths within 4 sigma,
  * any split of the photon range gives bit-identical tallies (fixed-point sums; the multi-GPU reduce is a plain sum),
  * scatter-to-primary ratio grows with slab thickness; the scatter image is smooth and centred.
"""
import numpy as np
import pytest

from deepdrr_b200 import Projector, Volume, geo, phantoms, scatter

pytestmark = pytest.mark.gpu


class _Dev:
    def __init__(self, carm, pose):
        self.source_to_detector_distance = carm.source_to_detector_distance
        self.camera_intrinsics = carm.camera_intrinsics
        self.detector_height, self.detector_width = carm.detector_height, carm.detector_width
        self._pose = pose

    def get_camera_projection(self):
        return self._pose


def _slab(thickness_mm, material="soft tissue", hu=0.0):
    n = 64
    nk = max(2, int(round(thickness_mm / 2.0)))
    data = np.full((n, n, nk), 1.0, dtype=np.float32)
    a = np.diag([4.0, 4.0, 2.0, 1.0])
    a[:3, 3] = [-4.0 * (n - 1) / 2, -4.0 * (n - 1) / 2, -2.0 * (nk - 1) / 2]
    return Volume(data, ({material: 0}, np.zeros(data.shape, np.uint16)), anatomical_from_IJK=geo.FrameTransform(a))


def _mono(e_kev):
    return np.array([[e_kev * 1000.0, 1.0], [e_kev * 1000.0 + 1.0, 0.0]])


def test_energy_bookkeeping_and_attenuation_law():
    L = 100.0
    v = _slab(L, "soft tissue")
    carm = phantoms.MobileCArmGeometry(sensor_width=64, sensor_height=64, pixel_size=0.5)   # narrow beam through the slab centre
    pose = carm.camera_projection(0.0, 0.0, (0, 0, 0))
    N = 2_000_000
    with Projector(v, device=_Dev(carm, pose), spectrum=_mono(60.0), neglog=False, scatter_num=N) as p:
        tally, c = scatter.simulate(p, pose, N, seed=7)
    emitted, missed, absorbed, prim, sc_det, sc_miss, n_ray, n_co = c
    assert missed == 0
    assert abs(emitted - (absorbed + prim + sc_det + sc_miss)) <= 1e-9 * emitted          # exact bookkeeping
    t = scatter.load_tables()
    names = [str(x) for x in t["names"]]
    m = names.index("soft tissue")
    ie = int(round((60000.0 - t["energy_eV"][0]) / (t["energy_eV"][1] - t["energy_eV"][0])))
    mfp_ray, mfp_co, mfp_ph, mfp_tot = t["mfp_mm"][m, ie, :4].astype(np.float64)
    # density 1.0 == nominal density of the table.  Nearly parallel rays: path = L / cos(theta) ~ L within 2e-4
    expect = np.exp(-L / mfp_tot)
    frac = prim / emitted
    sigma = np.sqrt(expect * (1 - expect) / N)
    assert abs(frac - expect) < 4 * sigma + 3e-4 * expect, (frac, expect, sigma)
    assert tally.sum() > 0 and abs(tally.sum() / 65536.0 - sc_det) <= 1e-6 * sc_det + 1.0
    # interaction-type ratio in a thin slab (photons that interact do so once, at 60 keV): Rayleigh / (Rayleigh + Compton)
    # must follow the inverse mean free paths; ~7e5 events -> sigma 3e-4, second interactions shift it by ~1e-3
    with Projector(_slab(10.0, "soft tissue"), device=_Dev(carm, pose), spectrum=_mono(60.0), neglog=False, scatter_num=N) as p:
        _, c2 = scatter.simulate(p, pose, 2 * N, seed=8)
    ratio = c2[6] / (c2[6] + c2[7])
    expect_r = (1 / mfp_ray) / (1 / mfp_ray + 1 / mfp_co)
    assert abs(ratio - expect_r) < 0.004, (ratio, expect_r)
    absorbed_frac = 1 - (c2[6] + c2[7]) / ((c2[6] + c2[7]) / (1 - mfp_tot / mfp_ph))       # photoelectric share of all events
    assert 0 < mfp_tot / mfp_ph < 0.2


def test_photon_range_splits_are_bit_identical():
    v = phantoms.thorax_volume((48, 48, 40), (8.5, 8.5, 10.0), seed=2)
    carm = phantoms.MobileCArmGeometry(sensor_width=96, sensor_height=64, pixel_size=3.0)
    pose = phantoms.c2_poses(1, seed=5, carm=carm)[0]
    N = 300_000
    with Projector(v, device=_Dev(carm, pose), spectrum="120KV_AL43", neglog=False, scatter_num=N) as p:
        whole, cw = scatter.simulate(p, pose, N, seed=3)
        parts = [scatter.simulate(p, pose, b - a, seed=3, photon_offset=a)[0] for a, b in ((0, 100_000), (100_000, 100_001), (100_001, N))]
        again, _ = scatter.simulate(p, pose, N, seed=3)
        other, _ = scatter.simulate(p, pose, N, seed=4)
    assert np.array_equal(whole, again)
    assert np.array_equal(whole, parts[0] + parts[1] + parts[2])          # what an 8-GPU all-reduce would sum
    assert not np.array_equal(whole, other)


def test_scatter_to_primary_ratio_and_projector_integration():
    carm = phantoms.MobileCArmGeometry(sensor_width=64, sensor_height=64, pixel_size=4.5)
    pose = carm.camera_projection(0.0, 0.0, (0, 0, 0))
    spr = []
    for L in (50.0, 150.0):
        v = _slab(L, "soft tissue")
        with Projector(v, device=_Dev(carm, pose), spectrum="90KV_AL40", neglog=False, scatter_num=3_000_000) as p:
            total = p.project()
            c = p.last_scatter_counters[0]
            p.scatter_num = 0
            primary = p.project()
        sc = total - primary
        assert np.all(sc >= 0) and sc.max() > 0
        centre, corner = sc[24:40, 24:40].mean(), sc[:8, :8].mean()
        assert centre > corner > 0                                        # smooth, centred scatter distribution
        spr.append(float(sc[24:40, 24:40].mean() / primary[24:40, 24:40].mean()))
        assert abs(c[0] - c[1:6].sum()) <= 1e-9 * c[0]
    assert 0.01 < spr[0] < spr[1] < 5.0                                   # more scatter behind a thicker slab
