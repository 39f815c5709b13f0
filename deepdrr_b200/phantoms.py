"""Deterministic synthetic phantoms and pose sets for BASELINE.json's configs.

No CT data exists offline, so every test / bench input is an analytic phantom built here from a
seed.  The recipes are the ones SURVEY.md section 8(d) fixes ("C1 input", "C2 input", "C3 input");
the C-arm camera follows the reference's ``MobileCArm`` (device/mobile_carm.py:73-84 defaults,
:223-258 transforms) and ``SimpleDevice`` style pinhole intrinsics (device/simple_device.py:66-78).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np

from . import geo
from .vol import Mesh, Volume


# ----------------------------------------------------------------------------------------------
# C1: 128^3 HU cylinder phantom, one 256^2 view
# ----------------------------------------------------------------------------------------------
def c1_hu(n: int = 128) -> np.ndarray:
    c = (n - 1) / 2.0
    i, j = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
    r = np.hypot(i - c, j - c)
    hu2 = np.where(r < 12 * n / 128, 1000.0, np.where(r < 48 * n / 128, 40.0, -1000.0))
    hu = np.repeat(hu2[:, :, None], n, axis=2).astype(np.float32)
    hu += np.random.default_rng(0).normal(0, 10, hu.shape).astype(np.float32)
    return hu


def c1_volume(n: int = 128) -> Volume:
    c = (n - 1) / 2.0
    a = np.array([[1, 0, 0, -c], [0, 1, 0, -c], [0, 0, 1, -c], [0, 0, 0, 1]], dtype=np.float64)
    return Volume.from_hu(c1_hu(n), anatomical_from_IJK=geo.FrameTransform(a))


def look_at_projection(source, direction, up, intrinsic: geo.CameraIntrinsicTransform) -> geo.CameraProjection:
    """z = viewing direction, y = -normalize(up - (up.z) z), x = y cross z, R = [x; y; z], t = -R s."""
    z = np.asarray(direction, dtype=np.float64)
    z = z / np.linalg.norm(z)
    up = np.asarray(up, dtype=np.float64)
    y = up - (up @ z) * z
    y = -y / np.linalg.norm(y)
    x = np.cross(y, z)
    r = np.stack([x, y, z], axis=0)
    t = -r @ np.asarray(source, dtype=np.float64)
    return geo.CameraProjection(intrinsic, geo.FrameTransform.from_rt(r, t))


def c1_camera(size: int = 256, direction=(0.3, 1.0, 0.2)) -> Tuple[geo.CameraProjection, float]:
    """W=H=size, pixel 1 mm (scaled so the field of view is fixed), SDD 1000, source 500 mm away."""
    pixel = 256.0 / size
    k = geo.CameraIntrinsicTransform.from_sizes((size, size), pixel, 1000.0)
    d = np.asarray(direction, dtype=np.float64)
    d = d / np.linalg.norm(d)
    proj = look_at_projection(-500.0 * d, d, (0, 0, 1), k)
    max_ray_length = math.sqrt(1000.0 ** 2 + 256.0 ** 2 + 256.0 ** 2)
    return proj, max_ray_length


# ----------------------------------------------------------------------------------------------
# C2: 512x512x400 synthetic thorax, MobileCArm geometry, random poses
# ----------------------------------------------------------------------------------------------
def thorax_hu(shape=(512, 512, 400), spacing=(0.8, 0.8, 1.0), noise_hu: float = 15.0, seed: int = 0) -> np.ndarray:
    """Analytic thorax: body ellipse, two lungs, spine, 12 rib rings, N(0, noise) HU noise.

    Axes: i = left-right (x), j = anterior-posterior (y), k = body axis (z); mm, centred on 0.
    Feature sizes scale with the physical extent so that reduced shapes keep the same anatomy.
    """
    ni, nj, nk = shape
    sx = spacing[0] * ni / 409.6  # 1.0 for the full-size config
    sy = spacing[1] * nj / 409.6
    sz = spacing[2] * nk / 400.0
    x = ((np.arange(ni) - (ni - 1) / 2.0) * spacing[0] / sx).astype(np.float32)[:, None, None]
    y = ((np.arange(nj) - (nj - 1) / 2.0) * spacing[1] / sy).astype(np.float32)[None, :, None]
    z = ((np.arange(nk) - (nk - 1) / 2.0) * spacing[2] / sz).astype(np.float32)[None, None, :]
    hu = np.full(shape, -1000.0, dtype=np.float32)
    body = (x / 180.0) ** 2 + (y / 130.0) ** 2
    hu[np.broadcast_to(body < 1.0, shape)] = 40.0
    for cx in (-75.0, 75.0):
        lung = ((x - cx) / 60.0) ** 2 + (y / 80.0) ** 2 + (z / 120.0) ** 2 < 1.0
        hu[lung] = -850.0
    re = np.sqrt((x / 165.0) ** 2 + (y / 115.0) ** 2)
    ring = (re > 0.93) & (re < 1.0)
    for n in range(12):
        zc = -165.0 + 30.0 * n
        rib = ring & (np.abs(z - zc) < 6.0)
        hu[rib] = 500.0
    spine = x ** 2 + (y - 70.0) ** 2 < 20.0 ** 2
    hu[np.broadcast_to(spine, shape)] = 700.0
    if noise_hu > 0:
        hu += np.random.default_rng(seed).normal(0, noise_hu, shape).astype(np.float32)
    return hu


# patient supine under a C-arm whose source sits below the table: anatomical x -> world x,
# anatomical z (body axis) -> world y, anatomical y (posterior) -> world -z.
SUPINE = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)


def thorax_volume(shape=(512, 512, 400), spacing=(0.8, 0.8, 1.0), noise_hu: float = 15.0, seed: int = 0) -> Volume:
    a = np.eye(4)
    for ax in range(3):
        a[ax, ax] = spacing[ax]
        a[ax, 3] = -spacing[ax] * (shape[ax] - 1) / 2.0
    return Volume.from_hu(thorax_hu(shape, spacing, noise_hu, seed),
                          anatomical_from_IJK=geo.FrameTransform(a),
                          world_from_anatomical=geo.FrameTransform(SUPINE))


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def _rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


class MobileCArmGeometry:
    """The camera model of the reference's ``MobileCArm`` (device/mobile_carm.py), pose only.

    camera3d_from_world = Rz(90 deg) . T(0, 0, source_to_isocenter) . inv(device_from_arm), with
    device_from_arm = (R = Ry(beta) Rx(alpha) [scipy ``from_euler("xy")``], t = isocenter)
    (device/mobile_carm.py:223-258; world_from_device = identity, gamma = 0).
    """

    def __init__(self, sensor_width=1536, sensor_height=1536, pixel_size=0.194,
                 source_to_detector_distance=1020.0, source_to_isocenter_vertical_distance=530.0):
        self.sensor_width = sensor_width
        self.sensor_height = sensor_height
        self.pixel_size = pixel_size
        self.source_to_detector_distance = source_to_detector_distance
        self.source_to_isocenter_vertical_distance = source_to_isocenter_vertical_distance
        self.camera_intrinsics = geo.CameraIntrinsicTransform.from_sizes(
            (sensor_width, sensor_height), pixel_size, source_to_detector_distance)

    @property
    def detector_width(self):
        return self.sensor_width * self.pixel_size

    @property
    def detector_height(self):
        return self.sensor_height * self.pixel_size

    @property
    def max_ray_length(self) -> float:
        """projector.py:641-647."""
        return math.sqrt(self.source_to_detector_distance ** 2 + self.detector_height ** 2 + self.detector_width ** 2)

    def camera_projection(self, alpha: float, beta: float, isocenter) -> geo.CameraProjection:
        rot = _rot_y(beta) @ _rot_x(alpha)
        device_from_arm = geo.FrameTransform.from_rt(rot, isocenter)
        cam_from_arm = geo.FrameTransform.from_rt(_rot_z(math.pi / 2)) @ geo.FrameTransform.from_translation(
            (0, 0, self.source_to_isocenter_vertical_distance))
        return geo.CameraProjection(self.camera_intrinsics, cam_from_arm @ device_from_arm.inv)


def c2_poses(n: int = 1000, seed: int = 1, carm: Optional[MobileCArmGeometry] = None,
             center=(0.0, 0.0, 0.0)) -> List[geo.CameraProjection]:
    """alpha, beta ~ U(-40, 40) deg; isocenter = centre + U(-30, 30)^3 mm."""
    carm = carm or MobileCArmGeometry()
    rng = np.random.default_rng(seed)
    ab = np.deg2rad(rng.uniform(-40, 40, size=(n, 2)))
    iso = np.asarray(center, dtype=np.float64) + rng.uniform(-30, 30, size=(n, 3))
    return [carm.camera_projection(ab[i, 0], ab[i, 1], iso[i]) for i in range(n)]


# ----------------------------------------------------------------------------------------------
# C3: K-wire tool volumes
# ----------------------------------------------------------------------------------------------
def kwire_volume(length_mm: float = 200.0, radius_mm: float = 1.0, tip_mm: float = 3.0,
                 spacing: float = 0.1, half_width: int = 10) -> Volume:
    """Synthetic K-wire: (21, 21, 2000) voxels at 0.1 mm, density 7.5 inside / 0 outside, every
    voxel labelled ``iron`` (reference: vol/kwire.py:90-102).  Tip at k = 0, wire along +k."""
    n = 2 * half_width + 1
    nk = int(round(length_mm / spacing))
    ij = (np.arange(n) - half_width) * spacing
    rr = np.hypot(ij[:, None], ij[None, :])[:, :, None]
    zk = (np.arange(nk) * spacing)[None, None, :]
    rad = np.where(zk < tip_mm, radius_mm * zk / tip_mm, radius_mm)
    data = np.where(rr <= rad, 7.5, 0.0).astype(np.float32)
    labels = np.zeros(data.shape, dtype=np.uint16)
    a = np.eye(4)
    a[0, 0] = a[1, 1] = a[2, 2] = spacing
    a[0, 3] = a[1, 3] = -half_width * spacing
    return Volume(data, ({"iron": 0}, labels), anatomical_from_IJK=geo.FrameTransform(a))


def place_kwire(wire: Volume, tip_world, direction_world) -> None:
    """Pose the wire so that its tip sits at ``tip_world`` and its axis points along ``direction``."""
    d = np.asarray(direction_world, dtype=np.float64)
    d = d / np.linalg.norm(d)
    helper = np.array([1.0, 0, 0]) if abs(d[0]) < 0.9 else np.array([0, 1.0, 0])
    x = np.cross(helper, d)
    x /= np.linalg.norm(x)
    y = np.cross(d, x)
    r = np.stack([x, y, d], axis=1)
    wire.world_from_anatomical = geo.FrameTransform.from_rt(r, tip_world)


def c3_scene(ct_shape=(512, 512, 400), ct_spacing=(0.8, 0.8, 1.0), ct: Optional[Volume] = None) -> List[Volume]:
    """CT + two K-wires crossing at 20 degrees near the CT centre; ``ct`` reuses an existing CT."""
    if ct is None:
        ct = thorax_volume(ct_shape, ct_spacing)
    w1, w2 = kwire_volume(), kwire_volume()
    half = math.radians(10.0)
    d1 = np.array([math.sin(half), math.cos(half), 0.2])
    d2 = np.array([-math.sin(half), math.cos(half), 0.2])
    place_kwire(w1, -100.0 * d1 / np.linalg.norm(d1) + np.array([0, 0, 5.0]), d1)
    place_kwire(w2, -100.0 * d2 / np.linalg.norm(d2) + np.array([0, 0, -5.0]), d2)
    return [ct, w1, w2]


def cone_poses(n: int, seed: int = 2, sensor: int = 384, pixel: float = 0.3, sdd: float = 1000.0,
               source_distance: float = 500.0, cone_deg: float = 30.0, axis=(0.0, 0.0, 1.0)) -> Tuple[List[geo.CameraProjection], float]:
    """n views whose direction lies within a ``cone_deg`` cone about ``axis`` (README.md:104-125 style
    sampling), pointed at the origin; sensor 384^2 at 0.3 mm, SDD 1000 (README.md:78-83)."""
    rng = np.random.default_rng(seed)
    k = geo.CameraIntrinsicTransform.from_sizes((sensor, sensor), pixel, sdd)
    axis = np.asarray(axis, dtype=np.float64)
    axis /= np.linalg.norm(axis)
    helper = np.array([1.0, 0, 0]) if abs(axis[0]) < 0.9 else np.array([0, 1.0, 0])
    e1 = np.cross(axis, helper)
    e1 /= np.linalg.norm(e1)
    e2 = np.cross(axis, e1)
    out = []
    cmin = math.cos(math.radians(cone_deg))
    for _ in range(n):
        ct = rng.uniform(cmin, 1.0)
        st = math.sqrt(1 - ct * ct)
        ph = rng.uniform(0, 2 * math.pi)
        d = ct * axis + st * (math.cos(ph) * e1 + math.sin(ph) * e2)
        out.append(look_at_projection(-source_distance * d, d, e1 if abs(d @ e1) < 0.9 else e2, k))
    return out, 4.0 * sdd


# ----------------------------------------------------------------------------------------------
# C4: procedural watertight meshes (the reference's STL fixtures are not shipped)
# ----------------------------------------------------------------------------------------------
def icosphere(radius: float = 10.0, subdivisions: int = 2):
    """(vertices [n,3], faces [m,3]) of a geodesic sphere, outward normals counter-clockwise."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return (np.array(v) * radius).astype(np.float32), np.array(f, dtype=np.int64)


def screw_mesh(length: float = 130.0, core_radius: float = 2.2, thread_radius: float = 3.25, pitch: float = 2.75,
               thread_length: float = 32.0, segments: int = 48, rings_per_mm: float = 4.0):
    """A watertight screw-like solid of revolution with a helical thread over the first ``thread_length`` mm
    (stand-in for data/6.5mmD_32mmThread_L130mm.STL: 6.5 mm thread diameter, 32 mm thread, 130 mm long).
    Axis = +z, tip at z = 0.  About 50 k triangles at the defaults."""
    nz = int(length * rings_per_mm) + 1
    z = np.linspace(0.0, length, nz)
    th = np.linspace(0.0, 2 * math.pi, segments, endpoint=False)
    zz, tt = np.meshgrid(z, th, indexing="ij")
    phase = (zz / pitch - tt / (2 * math.pi)) % 1.0
    tooth = np.clip(1.0 - np.abs(phase - 0.5) * 4.0, 0.0, 1.0)          # triangular thread profile
    taper = np.clip(zz / 6.0, 0.15, 1.0)                                  # pointed tip
    r = (core_radius + (thread_radius - core_radius) * tooth * (zz < thread_length)) * taper
    verts = np.stack([r * np.cos(tt), r * np.sin(tt), zz], axis=-1).reshape(-1, 3)
    faces = []
    for i in range(nz - 1):
        for j in range(segments):
            a, b = i * segments + j, i * segments + (j + 1) % segments
            c, d = a + segments, b + segments
            faces += [(a, b, d), (a, d, c)]
    bottom, top = len(verts), len(verts) + 1
    verts = np.concatenate([verts, [[0, 0, 0.0], [0, 0, length]]], axis=0)
    for j in range(segments):
        jn = (j + 1) % segments
        faces.append((bottom, jn, j))
        faces.append((top, (nz - 1) * segments + j, (nz - 1) * segments + jn))
    return verts.astype(np.float32), np.array(faces, dtype=np.int64)


def box_mesh(half=(5.0, 5.0, 5.0)):
    hx, hy, hz = half
    v = np.array([[-hx, -hy, -hz], [hx, -hy, -hz], [hx, hy, -hz], [-hx, hy, -hz], [-hx, -hy, hz], [hx, -hy, hz], [hx, hy, hz], [-hx, hy, hz]], dtype=np.float32)
    f = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6], [1, 2, 6], [1, 6, 5], [0, 4, 7], [0, 7, 3]], dtype=np.int64)
    return v, f
