"""The CPU restatement of the scatter transport (oracle/scatter_oracle.c): its own physics checks without
a GPU, and -- on the GPU -- the CUDA kernel against it on the SAME Philox subsequences.

The reference has no transport code (projector.py:530-531), so parity with it is unpinned; what these tests pin is that two
independent implementations of the published MC-GPU scheme (csrc/drr_scatter.cu and the plain-C restatement), fed the same
tables, poses and per-photon random streams, agree far inside the Monte Carlo noise: most photon histories are identical, the
rest differ where libm and the CUDA math library round a logarithm or a sine differently at a decision threshold.
"""
import numpy as np
import pytest

from deepdrr_b200 import Projector, Volume, geo, phantoms, scatter
from deepdrr_b200.spectral_data import get_spectrum, spectrum_tables
from oracle import scatter_oracle


class _Dev:
    def __init__(self, carm, pose):
        self.source_to_detector_distance = carm.source_to_detector_distance
        self.camera_intrinsics = carm.camera_intrinsics
        self.detector_height, self.detector_width = carm.detector_height, carm.detector_width
        self._pose = pose

    def get_camera_projection(self):
        return self._pose


def _slab(thickness_mm, material="soft tissue"):
    n = 64
    nk = max(2, int(round(thickness_mm / 2.0)))
    data = np.full((n, n, nk), 1.0, dtype=np.float32)
    a = np.diag([4.0, 4.0, 2.0, 1.0])
    a[:3, 3] = [-4.0 * (n - 1) / 2, -4.0 * (n - 1) / 2, -2.0 * (nk - 1) / 2]
    return Volume(data, ({material: 0}, np.zeros(data.shape, np.uint16)), anatomical_from_IJK=geo.FrameTransform(a))


def _mono(e_kev):
    return np.array([[e_kev * 1000.0, 1.0], [e_kev * 1000.0 + 1.0, 0.0]])


def test_restatement_obeys_the_attenuation_law_and_conserves_energy():
    L = 100.0
    v = _slab(L)
    carm = phantoms.MobileCArmGeometry(sensor_width=64, sensor_height=64, pixel_size=0.5)
    pose = carm.camera_projection(0.0, 0.0, (0, 0, 0))
    e, pdf = spectrum_tables(get_spectrum(_mono(60.0)))
    N = 150_000
    tally, c = scatter_oracle.simulate(v, ["soft tissue"], e, pdf, pose, carm.source_to_detector_distance, N, seed=7)
    emitted, missed, absorbed, prim, sc_det, sc_miss, n_ray, n_co = c
    assert missed == 0 and abs(emitted - (absorbed + prim + sc_det + sc_miss)) <= 1e-9 * emitted
    t = scatter.load_tables()
    m = [str(x) for x in t["names"]].index("soft tissue")
    ie = int(round((60000.0 - t["energy_eV"][0]) / (t["energy_eV"][1] - t["energy_eV"][0])))
    mfp_ray, mfp_co, mfp_ph, mfp_tot = t["mfp_mm"][m, ie, :4].astype(np.float64)
    expect = np.exp(-L / mfp_tot)
    sigma = np.sqrt(expect * (1 - expect) / N)
    assert abs(prim / emitted - expect) < 4 * sigma + 3e-4 * expect
    assert abs(tally.sum() / 65536.0 - sc_det) <= 1e-6 * sc_det + 1.0
    # the same photon ids in two pieces give the same tally (the property the multi-GPU split relies on)
    t1, _ = scatter_oracle.simulate(v, ["soft tissue"], e, pdf, pose, carm.source_to_detector_distance, 50_000, seed=7)
    t2, _ = scatter_oracle.simulate(v, ["soft tissue"], e, pdf, pose, carm.source_to_detector_distance, N - 50_000, seed=7, photon_offset=50_000)
    assert np.array_equal(t1 + t2, tally)


def test_compton_sampler_binding_and_doppler_broadening():
    """The impulse-approximation sampler on its own (water, 60 keV, 200 000 events) against what the model must show:
    * angular distribution = Klein-Nishina x incoherent scattering function: forward scattering is suppressed, so the mean
      energy loss per event lies a few per cent ABOVE the free-electron Klein-Nishina value (5.6 keV at 60 keV);
    * at a fixed angle the scattered energy is spread around the Compton line (Doppler broadening by the electrons' momentum
      distribution): a fraction of a per cent to a few per cent at 60 keV, and zero in a sampler without it;
    * no photon gains more than the binding-free kinematics allow, none falls below zero."""
    E = 60000.0
    cost, e_out = scatter_oracle.compton_samples("water", E, 200_000, seed=5)
    k = E / 510998.918
    mu = np.linspace(-1, 1, 20001)
    tau = 1 / (1 + k * (1 - mu))
    kn = tau ** 2 * (tau + 1 / tau - (1 - mu ** 2))
    kn_loss = E * np.trapezoid((1 - tau) * kn, mu) / np.trapezoid(kn, mu)
    loss = float(np.mean(E - e_out.astype(np.float64)))
    assert 1.0 < loss / kn_loss < 1.15, (loss, kn_loss)
    kn_cos = np.trapezoid(mu * kn, mu) / np.trapezoid(kn, mu)
    assert float(cost.mean()) < kn_cos                                   # fewer forward events than for free electrons
    line = E / (1 + k * (1 - cost.astype(np.float64)))                   # Compton line at each sampled angle
    rel = (e_out - line) / line
    back = cost < -0.5                                                    # backscatter: largest momentum transfer, widest line
    assert 0.002 < float(np.std(rel[back])) < 0.05, float(np.std(rel[back]))
    assert abs(float(np.mean(rel[back]))) < 0.01                          # ... and centred on the line
    assert e_out.min() > 0 and e_out.max() < 1.02 * E


@pytest.mark.gpu
def test_cuda_kernel_matches_the_restatement_on_the_same_photon_streams():
    v = phantoms.thorax_volume((48, 48, 40), (8.5, 8.5, 10.0), seed=2)
    carm = phantoms.MobileCArmGeometry(sensor_width=96, sensor_height=64, pixel_size=3.0)
    pose = phantoms.c2_poses(1, seed=5, carm=carm)[0]
    N = 100_000
    with Projector(v, device=_Dev(carm, pose), spectrum="120KV_AL43", neglog=False, scatter_num=N) as p:
        mats = p.all_materials
        g_tally, g_c = scatter.simulate(p, pose, N, seed=3, photon_offset=12345)
    e, pdf = spectrum_tables(get_spectrum("120KV_AL43"))
    c_tally, c_c = scatter_oracle.simulate(v, mats, e, pdf, pose, carm.source_to_detector_distance, N, seed=3, photon_offset=12345)
    # the source (spectrum CDF, direction, weight) is plain arithmetic on the random numbers: identical streams -> same sums
    assert abs(g_c[0] - c_c[0]) <= 1e-6 * c_c[0], "emitted energy differs: the Philox restatement is off"
    assert abs(g_c[1] - c_c[1]) <= 1e-6 * c_c[0]
    for k in (6, 7):                                                  # number of Rayleigh / Compton events
        assert abs(g_c[k] - c_c[k]) <= 0.003 * c_c[k], (k, g_c[k], c_c[k])
    for k in (2, 3, 4, 5):                                            # absorbed / primary / scattered-detected / scattered-lost energy
        assert abs(g_c[k] - c_c[k]) <= 0.005 * c_c[k], (k, g_c[k], c_c[k])
    same = float(np.mean(g_tally == c_tally))
    assert same > 0.85, f"only {same:.2%} of the pixels carry identical tallies"
    # coarse 8x8 bins, 3 sigma: the variance of a bin is the sum of the squared deposits, ~ bin total x mean deposit
    G = g_tally.astype(np.float64).reshape(8, 8, 12, 8).sum(axis=(1, 3))
    C = c_tally.astype(np.float64).reshape(8, 8, 12, 8).sum(axis=(1, 3))
    hits = max(1, int((c_tally > 0).sum()))
    mean_dep = c_tally.sum() / hits
    sigma = np.sqrt(2.0 * np.maximum(C, mean_dep) * mean_dep)
    assert np.all(np.abs(G - C) <= 3.0 * sigma), float((np.abs(G - C) / sigma).max())


def _two_slabs():
    """Two 64 x 64 slabs of soft tissue, 40 mm thick each, 60 mm of vacuum between them, plus a bone plate that overlaps the lower
    slab and outranks it (priority 0 < 1): the union a multi-volume scene has to track through."""
    lower, upper = _slab(40.0), _slab(40.0)
    lower.translate((0.0, 0.0, -50.0))
    upper.translate((0.0, 0.0, 50.0))
    n = 24
    data = np.full((n, n, 6), 1.9, dtype=np.float32)
    a = np.diag([4.0, 4.0, 2.0, 1.0])
    a[:3, 3] = [-4.0 * (n - 1) / 2, -4.0 * (n - 1) / 2, -5.0]
    plate = Volume(data, ({"bone": 0}, np.zeros(data.shape, np.uint16)), anatomical_from_IJK=geo.FrameTransform(a))
    plate.translate((0.0, 0.0, -50.0))
    return [lower, upper, plate], [1, 1, 0]


def test_restatement_tracks_through_several_volumes():
    """Unscattered fraction through two separated slabs = exp(-(L1 + L2) / mfp): the vacuum between them attenuates nothing, and a
    scene split into two volumes behaves like the same matter in one."""
    lower, upper = _slab(40.0), _slab(40.0)
    lower.translate((0.0, 0.0, -50.0))
    upper.translate((0.0, 0.0, 50.0))
    carm = phantoms.MobileCArmGeometry(sensor_width=64, sensor_height=64, pixel_size=0.5)
    pose = carm.camera_projection(0.0, 0.0, (0, 0, 0))
    e, pdf = spectrum_tables(get_spectrum(_mono(60.0)))
    N = 150_000
    _, c = scatter_oracle.simulate([lower, upper], ["soft tissue"], e, pdf, pose, carm.source_to_detector_distance, N, seed=2)
    t = scatter.load_tables()
    m = [str(x) for x in t["names"]].index("soft tissue")
    ie = int(round((60000.0 - t["energy_eV"][0]) / (t["energy_eV"][1] - t["energy_eV"][0])))
    mfp_tot = float(t["mfp_mm"][m, ie, 3])
    expect = np.exp(-80.0 / mfp_tot)
    assert abs(c[3] / c[0] - expect) < 4 * np.sqrt(expect * (1 - expect) / N) + 3e-4 * expect
    assert abs(c[0] - c[1:6].sum()) <= 1e-9 * c[0]


@pytest.mark.gpu
def test_cuda_multi_volume_scatter_matches_the_restatement():
    volumes, priorities = _two_slabs()
    carm = phantoms.MobileCArmGeometry(sensor_width=96, sensor_height=64, pixel_size=3.0)
    pose = carm.camera_projection(0.15, -0.1, (5.0, 0.0, 0.0))
    N = 100_000
    with Projector(volumes, priorities=priorities, device=_Dev(carm, pose), spectrum="90KV_AL40", neglog=False, scatter_num=N) as p:
        mats = p.all_materials
        g_tally, g_c = scatter.simulate(p, pose, N, seed=9)
        img = p.project()                                               # primary + scatter through the public call
        p.scatter_num = 0
        primary = p.project()
    e, pdf = spectrum_tables(get_spectrum("90KV_AL40"))
    c_tally, c_c = scatter_oracle.simulate(volumes, mats, e, pdf, pose, carm.source_to_detector_distance, N, seed=9, priorities=priorities)
    assert abs(g_c[0] - c_c[0]) <= 1e-6 * c_c[0]
    for k in (2, 3, 4, 5, 6, 7):
        assert abs(g_c[k] - c_c[k]) <= 0.006 * c_c[k], (k, g_c[k], c_c[k])
    assert float(np.mean(g_tally == c_tally)) > 0.85
    assert np.all(img >= primary) and float((img - primary).sum()) > 0
    # the bone plate outranks the slab it sits in: more photoabsorption than the same scene without it
    with Projector(volumes[:2], device=_Dev(carm, pose), spectrum="90KV_AL40", neglog=False, scatter_num=N) as p:
        _, c2 = scatter.simulate(p, pose, N, seed=9)
    assert g_c[2] > 1.02 * c2[2]
