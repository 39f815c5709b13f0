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
