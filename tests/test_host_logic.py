"""Host-side logic that mirrors the reference's Projector / scene bookkeeping (no GPU needed)."""
import numpy as np
import pytest

import cases
from deepdrr_b200 import Projector, Volume, geo, phantoms, vol
from deepdrr_b200.parallel import shard_range, shard_sizes
from deepdrr_b200.projector import DeprecationError
from deepdrr_b200.scene import default_priorities, material_universe, remap_labels


def _tiny_volume(materials=("air", "soft tissue", "bone")):
    hu = np.zeros((4, 5, 6), dtype=np.float32)
    hu[1:3] = 500.0
    hu[3] = -1000.0
    return Volume.from_hu(hu)


def test_threshold_segmentation_and_density():
    v = _tiny_volume()
    assert v.materials[0] == {"air": 0, "soft tissue": 1, "bone": 2}          # load_dicom.py:132-143 dict order
    assert v.materials[1].dtype == np.uint16 and v.data.dtype == np.float32
    assert set(np.unique(v.materials[1])) == {0, 1, 2}
    hu = np.array([-1000.0, 0.0, 1000.0])
    d = vol.convert_hounsfield_to_density(hu.copy())
    assert np.allclose(d, [0.001, 1.03, 1.6186], atol=1e-12)                   # vol/volume.py:338-351


def test_material_universe_sorted_and_label_remap():
    v = _tiny_volume()
    mats = material_universe([v])
    assert mats == ["air", "bone", "soft tissue"]                              # projector.py:547-559
    lab = remap_labels(v, mats)
    assert lab.dtype == np.uint8
    # dict order air, soft tissue, bone -> sorted index 0, 2, 1 (projector.py:1499-1509)
    assert np.array_equal(lab, np.array([0, 2, 1], dtype=np.uint8)[v.materials[1]])
    assert material_universe([v], attenuate_outside_volume=True) == ["air", "bone", "soft tissue"]
    assert default_priorities(3) == [2, 1, 0]                                  # projector.py:489-492


def test_format_materials_later_masks_win():
    a = np.zeros((2, 2, 2), bool); a[0] = True
    b = np.zeros((2, 2, 2), bool); b[0, 0] = True
    d, lab = vol.format_materials({"x": a, "y": b})
    assert d == {"x": 0, "y": 1}
    assert lab[0, 0, 0] == 1 and lab[0, 1, 0] == 0 and lab[1, 0, 0] == 0      # unlabeled voxels stay 0


def test_pose_arrays_follow_reference_formulas():
    v = phantoms.c1_volume(16)
    proj, _ = phantoms.c1_camera(32)
    w2i, src, ijk = geo.pose_arrays(proj, [v])
    k = proj.intrinsic.data
    r = proj.camera3d_from_world.data[:3, :3]
    assert np.allclose(w2i.reshape(3, 3), (r.T @ np.linalg.inv(k)).astype(np.float32))
    c = proj.center_in_world
    assert np.allclose(np.linalg.norm(c), 500.0)
    m = np.linalg.inv(v.world_from_IJK.data)
    assert np.allclose(src[0], (m[:3, :3] @ c + m[:3, 3]).astype(np.float32))
    assert np.allclose(ijk[0], m[:3, :].astype(np.float32).reshape(12))
    # the ray through the principal point is the viewing direction
    d = w2i.reshape(3, 3) @ np.array([16.0, 16.0, 1.0])
    assert np.allclose(d / np.linalg.norm(d), r[2], atol=1e-6)


def test_mobile_carm_geometry():
    carm = phantoms.MobileCArmGeometry()
    assert carm.camera_intrinsics.sensor_size == (1536, 1536)
    assert abs(carm.max_ray_length - 1103.6) < 0.1
    p = carm.camera_projection(0.0, 0.0, (0, 0, 0))
    assert np.allclose(p.center_in_world, [0, 0, -530.0], atol=1e-9)           # source below the isocenter
    d = p.world_from_index[:3] @ np.array([768.0, 768.0, 1.0])
    assert np.allclose(d / np.linalg.norm(d), [0, 0, 1.0], atol=1e-9)


def test_projector_constructor_contract():
    v = _tiny_volume()
    k = geo.CameraIntrinsicTransform.from_sizes((8, 8), 1.0, 100.0)
    p = Projector(v, camera_intrinsics=k)
    assert p.priorities == [0] and p.all_materials == ["air", "bone", "soft tissue"] and p.step == 0.1
    assert p.volume is v and p.camera_intrinsics is k and p.source_to_detector_distance == -1
    with pytest.raises(RuntimeError):
        p.project(phantoms.c1_camera(8)[0])                                    # not initialized (projector.py:629-630)
    with pytest.raises(KeyError):
        Projector(v, spectrum="nope")
    with pytest.raises(TypeError):
        Projector(v, spectrum=12)
    with pytest.raises(ValueError):
        Projector(v, scatter_num=-1)
    with pytest.raises(ValueError):
        Projector(v, scatter_num=10)                                           # needs a device (projector.py:527-528)
    with pytest.raises(ValueError):
        Projector(v, max_mesh_hits=6)
    with pytest.raises(ValueError):
        Projector(v, add_scatter=True, scatter_num=5)
    with pytest.raises(ValueError):
        Projector(["not a volume"])
    bad = Volume(np.zeros((2, 2, 2), np.float32), ({"unobtainium": 0}, np.zeros((2, 2, 2), np.uint16)))
    with pytest.raises(ValueError):
        Projector(bad)
    with pytest.raises(AttributeError):
        Projector([v, v]).volume
    with pytest.raises(AssertionError):
        Projector([v, v], priorities=[0, 5])
    with pytest.raises(DeprecationError):
        Projector(v, camera_intrinsics=k).project_over_carm_range()

    class Dev:
        source_to_detector_distance = 1000.0
        camera_intrinsics = k
    # the reference raises DeprecationError here (projector.py:530-531); scatter is functional again in this build
    assert Projector(v, device=Dev(), scatter_num=100).scatter_num == 100


def test_initialize_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = Projector(_tiny_volume(), camera_intrinsics=geo.CameraIntrinsicTransform.from_sizes((8, 8), 1.0, 100.0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.initialize()
    assert not p.initialized


def test_shard_ranges_cover_everything_once():
    for n in (0, 1, 7, 8, 1000, 10001):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = shard_sizes(n, w)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_devices_follow_reference_geometry():
    from deepdrr_b200 import device

    c = device.MobileCArm(alpha=10, beta=-20, isocenter=(5, 6, 7))
    g = phantoms.MobileCArmGeometry().camera_projection(np.radians(10), np.radians(-20), (5, 6, 7))
    assert np.allclose(c.get_camera_projection().camera3d_from_world.data, g.camera3d_from_world.data, atol=1e-12)
    assert c.camera_intrinsics.sensor_size == (1536, 1536) and abs(c.detector_width - 297.984) < 1e-9
    c.move_by(delta_alpha=500)                                   # clipped to max_alpha (mobile_carm.py:295-303)
    assert abs(np.degrees(c.alpha) - 110) < 1e-9
    c.move_to(alpha=0, beta=0, isocenter_in_world=(1, 2, 3))
    assert np.allclose(c.isocenter, (1, 2, 3)) and np.allclose(c.source_in_world, (1, 2, 3 - 530.0))
    batch = c.camera_projections([10, 20], [-20, 5], np.array([[5, 6, 7], [0, 0, 0]]))
    assert np.allclose(batch[0].camera3d_from_world.data, g.camera3d_from_world.data, atol=1e-12)
    with pytest.raises(ValueError):
        device.MobileCArm(isocenter=(1000, 0, 0), enforce_isocenter_bounds=True)

    s = device.SimpleDevice(sensor_height=64, sensor_width=80, pixel_size=2.0)
    s.set_view([10, 20, 30], [0.3, 1.0, 0.2], [0, 0, 1])
    p = s.get_camera_projection()
    d = p.world_from_index[:3] @ np.array([40, 32, 1.0])
    d /= np.linalg.norm(d)
    assert np.allclose(d, np.array([0.3, 1, 0.2]) / np.linalg.norm([0.3, 1, 0.2]), atol=1e-9)
    assert np.allclose(p.center_in_world + 500 * d, [10, 20, 30], atol=1e-9)       # the point is mid-way (fraction 0.5)
    up_cam = p.camera3d_from_world.R @ np.array([0, 0, 1.0])
    assert up_cam[1] < -0.9 and abs(up_cam[0]) < 1e-9                              # world up shows as -y (image up)
    assert s.camera_intrinsics.sensor_size == (80, 64) and s.camera_intrinsics.fx == 500.0


def test_hu_volume_is_lazy_and_equivalent_on_the_host():
    from deepdrr_b200 import HUVolume

    hu = phantoms.c1_hu(12)
    a, b = HUVolume(hu), Volume.from_hu(hu)
    assert a.materials[0] == b.materials[0] and a._host is None and a.shape == b.shape
    assert Projector(a, camera_intrinsics=geo.CameraIntrinsicTransform.from_sizes((8, 8), 1.0, 100.0)).all_materials == ["air", "bone", "soft tissue"]
    assert a._host is None                                                          # nothing materialised so far
    assert np.array_equal(a.data, b.data) and np.array_equal(np.asarray(a.materials[1]), b.materials[1])


def test_carm_poses_against_matrices_derived_from_the_reference_formulas():
    """device/mobile_carm.py:223-258, written out with scipy exactly as the reference writes it (independent of this package's
    code) and by hand for the poses where the matrices are obvious; the batched generator must reproduce both, for every view,
    and hand the Projector the same kernel arrays as the per-view path."""
    from scipy.spatial.transform import Rotation

    from deepdrr_b200 import device

    def reference(alpha, beta, gamma, iso, vertical, horizontal, left, world_from_device):
        def rt(r=None, t=None):
            m = np.eye(4)
            if r is not None:
                m[:3, :3] = r
            if t is not None:
                m[:3, 3] = t
            return m
        device_from_arm = rt(Rotation.from_euler("xy", [alpha, beta]).as_matrix(), iso)                 # :223-227
        camera3d_from_arm = rt(t=[0, -horizontal, vertical])                                            # :239-245
        if left:
            camera3d_from_arm = rt(Rotation.from_euler("z", 90, degrees=True).as_matrix()) @ camera3d_from_arm   # :246-250
        gamma_rotation = rt(Rotation.from_euler("z", gamma, degrees=False).as_matrix())                 # :255-257
        return gamma_rotation @ camera3d_from_arm @ np.linalg.inv(device_from_arm) @ np.linalg.inv(world_from_device)  # :259, :276

    wfd = geo.FrameTransform.from_rt(Rotation.from_euler("zyx", [0.3, -0.2, 0.1]).as_matrix(), (12.0, -7.0, 3.0))
    c = device.MobileCArm(world_from_device=wfd, gamma=0.15, degrees=False, source_to_isocenter_horizontal_offset=4.0)
    rng = np.random.default_rng(3)
    n = 300
    al, be = rng.uniform(-1.5, 1.5, n), rng.uniform(-1.5, 1.5, n)
    iso = rng.uniform(-80, 80, (n, 3))
    batch = c.camera3d_from_world_batch(al, be, iso, degrees=False)
    assert batch.shape == (n, 4, 4)
    for i in range(n):
        want = reference(al[i], be[i], 0.15, iso[i], 530.0, 4.0, True, wfd.data)
        assert np.allclose(batch[i], want, atol=1e-9), i
    # by hand: no rotation, isocentre at the origin -> the source sits 530 mm below, the camera is turned 90 degrees about z
    plain = device.MobileCArm()
    m0 = plain.camera3d_from_world_batch([0.0], [0.0], None)[0]
    assert np.allclose(m0, [[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 530.0], [0, 0, 0, 1]], atol=1e-12)
    # alpha = 90 degrees about x: arm_from_device = Rx(-90): (x, y, z) -> (x, z, -y); then + (0, 0, 530), then Rz(90)
    m1 = plain.camera3d_from_world_batch([90.0], [0.0], None)[0]
    assert np.allclose(m1, [[0, 0, -1, 0], [1, 0, 0, 0], [0, -1, 0, 530.0], [0, 0, 0, 1]], atol=1e-12)
    # beta = 90 degrees about y with the isocentre at (10, 0, 0): arm = Ry(-90) (p - iso): (x, y, z) -> (-z, y, x - 10)
    m2 = plain.camera3d_from_world_batch([0.0], [90.0], np.array([[10.0, 0, 0]]))[0]
    assert np.allclose(m2, [[0, -1, 0, 0], [0, 0, -1, 0], [1, 0, 0, 520.0], [0, 0, 0, 1]], atol=1e-12)
    # the objects built from the batch equal the stateful per-view path, and so do the kernel arrays the Projector uploads
    projs = c.camera_projections(al[:5], be[:5], iso[:5], degrees=False)
    vol_ = phantoms.thorax_volume((8, 8, 6), (40.0, 40.0, 60.0))
    w_b, s_b, a_b = geo.pose_arrays_batch(projs, [vol_])
    for i in range(5):
        c.move_to(isocenter=iso[i], alpha=al[i], beta=be[i], degrees=False)
        one = c.get_camera_projection()
        assert np.allclose(one.camera3d_from_world.data, projs[i].camera3d_from_world.data, atol=1e-9)
        w1, s1, a1 = geo.pose_arrays(projs[i], [vol_])
        assert np.array_equal(w1, w_b[i]) and np.array_equal(s1, s_b[i]) and np.array_equal(a1, a_b[i])
