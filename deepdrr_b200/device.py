"""Imaging devices that produce camera projections for the Projector.

Pose-generating subset of the reference's ``deepdrr.device`` package, without killeengeo / pyvista:

* ``Device``        -- the interface ``Projector`` uses (device/device.py:7-209): ``camera_intrinsics``,
                       ``source_to_detector_distance``, ``detector_width/height``, ``get_camera_projection()``.
* ``SimpleDevice``  -- point / direction / up interface (device/simple_device.py:10-178).
* ``MobileCArm``    -- Cios-Fusion-like C-arm with alpha / beta / isocenter (device/mobile_carm.py:58-420).

Extra: ``MobileCArm.camera_projections(alphas, betas, isocenters)`` builds a whole batch of poses at once,
which is what the batched ``Projector.project(*poses)`` wants.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from . import geo


def _radians(x, degrees):
    return math.radians(x) if degrees else float(x)


def _rotvec_matrix(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(axis)
    if n < 1e-12 or abs(angle) < 1e-15:
        return np.eye(3)
    k = axis / n
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)


def _angle(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    c = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
    return math.acos(max(-1.0, min(1.0, c)))


class Device:
    """device/device.py:7-209, pose part."""

    sensor_height: int
    sensor_width: int
    pixel_size: float
    source_to_detector_distance: float
    world_from_device: geo.FrameTransform

    @property
    def device_from_world(self) -> geo.FrameTransform:
        return self.world_from_device.inv

    @property
    def camera_intrinsics(self) -> geo.CameraIntrinsicTransform:
        return geo.CameraIntrinsicTransform.from_sizes((self.sensor_width, self.sensor_height), self.pixel_size,
                                                       self.source_to_detector_distance)

    @property
    def detector_height(self) -> float:
        return self.sensor_height * self.pixel_size

    @property
    def detector_width(self) -> float:
        return self.sensor_width * self.pixel_size

    @property
    def camera3d_from_world(self) -> geo.FrameTransform:
        return self.device_from_camera3d.inv @ self.device_from_world

    def get_camera_projection(self) -> geo.CameraProjection:
        return geo.CameraProjection(self.camera_intrinsics, self.camera3d_from_world)

    @property
    def index_from_world(self) -> geo.CameraProjection:
        return self.get_camera_projection()

    @property
    def principle_ray_in_world(self) -> np.ndarray:
        return (self.world_from_device @ self.device_from_camera3d).R @ np.array([0.0, 0.0, 1.0])

    @property
    def source_in_world(self) -> np.ndarray:
        return self.get_camera_projection().center_in_world


class SimpleDevice(Device):
    """device/simple_device.py:10-178."""

    def __init__(self, sensor_height: int = 384, sensor_width: int = 384, pixel_size: float = 1.0,
                 source_to_detector_distance: float = 1000.0, world_from_device=None):
        self.sensor_height, self.sensor_width, self.pixel_size = sensor_height, sensor_width, pixel_size
        self.source_to_detector_distance = source_to_detector_distance
        self.world_from_device = geo.frame_transform(world_from_device)
        self._device_from_camera3d = geo.FrameTransform.identity()
        self.set_view([0, 0, 0], [0, 0, 1], [0, -1, 0])

    def scale_sensor(self, detector_size: float):
        if self.sensor_height < self.sensor_width:
            self.pixel_size = detector_size / self.sensor_height
        else:
            self.pixel_size = detector_size / self.sensor_width

    def set_view(self, point=None, direction=None, up=None, source_to_point_distance: Optional[float] = None,
                 source_to_point_fraction: float = 0.5):
        """Same construction as simple_device.py:80-165: ray frame (z along the direction), then a roll about z that
        brings the projected up-vector onto -y, then the camera pulled back along -z."""
        if source_to_point_distance is None:
            source_to_point_distance = self.source_to_detector_distance * source_to_point_fraction
        dfw = self.device_from_world
        if point is None:
            point_in_device = self._device_from_camera3d @ np.array([0.0, 0.0, source_to_point_distance])
        else:
            point_in_device = dfw @ np.asarray(point, dtype=np.float64)
        if direction is None:
            dir_in_device = self._device_from_camera3d.R @ np.array([0.0, 0.0, 1.0])
        else:
            dir_in_device = dfw.R @ np.asarray(direction, dtype=np.float64)
        up_in_device = np.array([0.0, -1.0, 0.0]) if up is None else dfw.R @ np.asarray(up, dtype=np.float64)
        z = np.array([0.0, 0.0, 1.0])
        rv = np.cross(z, dir_in_device)
        if np.linalg.norm(rv) < 1e-6:
            rot = _rotvec_matrix([1, 0, 0], math.pi) if z @ dir_in_device < 0 else np.eye(3)
        else:
            rot = _rotvec_matrix(rv, _angle(z, dir_in_device))
        device_from_ray = geo.FrameTransform.from_rt(rot, point_in_device)
        up_in_ray = device_from_ray.inv.R @ up_in_device
        up_plane = np.array([up_in_ray[0], up_in_ray[1], 0.0])
        neg_y = np.array([0.0, -1.0, 0.0])
        rv = np.cross(neg_y, up_plane)
        rot2 = np.eye(3) if np.linalg.norm(rv) < 1e-6 else _rotvec_matrix(rv, _angle(neg_y, up_plane))
        ray_from_ray_up = geo.FrameTransform.from_rt(rot2)
        ray_up_from_camera3d = geo.FrameTransform.from_translation((0, 0, -source_to_point_distance))
        self._device_from_camera3d = device_from_ray @ ray_from_ray_up @ ray_up_from_camera3d

    @property
    def device_from_camera3d(self) -> geo.FrameTransform:
        return self._device_from_camera3d


class MobileCArm(Device):
    """device/mobile_carm.py:58-420 (pose part; bounds enforcement as in :295-303)."""

    def __init__(self, world_from_device=None, isocenter=(0, 0, 0), alpha: float = 0, beta: float = 0, gamma: float = 0,
                 degrees: bool = True, horizontal_movement: float = 200, vertical_travel: float = 430, min_alpha: float = -40,
                 max_alpha: float = 110, min_beta: float = -225, max_beta: float = 225, source_to_detector_distance: float = 1020,
                 source_to_isocenter_vertical_distance: float = 530, source_to_isocenter_horizontal_offset: float = 0,
                 sensor_height: int = 1536, sensor_width: int = 1536, pixel_size: float = 0.194, rotate_camera_left: bool = True,
                 enforce_isocenter_bounds: bool = False):
        self.world_from_device = geo.frame_transform(world_from_device)
        self.enforce_isocenter_bounds = enforce_isocenter_bounds
        self.isocenter = np.asarray(isocenter, dtype=np.float64).reshape(3).copy()
        self.alpha, self.beta, self.gamma = _radians(alpha, degrees), _radians(beta, degrees), _radians(gamma, degrees)
        self.horizontal_movement, self.vertical_travel = horizontal_movement, vertical_travel
        self.min_alpha, self.max_alpha = _radians(min_alpha, degrees), _radians(max_alpha, degrees)
        self.min_beta, self.max_beta = _radians(min_beta, degrees), _radians(max_beta, degrees)
        self.source_to_detector_distance = source_to_detector_distance
        self.source_to_isocenter_vertical_distance = source_to_isocenter_vertical_distance
        self.source_to_isocenter_horizontal_offset = source_to_isocenter_horizontal_offset
        self.sensor_height, self.sensor_width, self.pixel_size = sensor_height, sensor_width, pixel_size
        self.rotate_camera_left = rotate_camera_left
        if enforce_isocenter_bounds and (np.any(self.isocenter < self.min_isocenter) or np.any(self.isocenter > self.max_isocenter)):
            raise ValueError(f"isocenter {self.isocenter} is out of bounds. Use world_from_device transform to position the carm in the world.")
        self._enforce_bounds()

    @property
    def max_isocenter(self) -> np.ndarray:
        return np.array([self.horizontal_movement, self.horizontal_movement, self.vertical_travel]) / 2

    @property
    def min_isocenter(self) -> np.ndarray:
        return -self.max_isocenter

    def _enforce_bounds(self):
        if self.enforce_isocenter_bounds:
            self.isocenter = np.clip(self.isocenter, self.min_isocenter, self.max_isocenter)
        self.alpha = float(np.clip(self.alpha, self.min_alpha, self.max_alpha))
        self.beta = float(np.clip(self.beta, self.min_beta, self.max_beta))

    @staticmethod
    def _camera3d_from_device(alpha, beta, gamma, isocenter, vertical, horizontal, rotate_left) -> np.ndarray:
        ca, sa, cb, sb = math.cos(alpha), math.sin(alpha), math.cos(beta), math.sin(beta)
        rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
        ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
        device_from_arm = np.eye(4)
        device_from_arm[:3, :3] = ry @ rx                     # scipy Rotation.from_euler("xy", [alpha, beta])
        device_from_arm[:3, 3] = isocenter
        cam_from_arm = np.eye(4)
        cam_from_arm[:3, 3] = [0, -horizontal, vertical]
        if rotate_left:
            rz = np.eye(4)
            rz[:3, :3] = [[0, -1, 0], [1, 0, 0], [0, 0, 1]]
            cam_from_arm = rz @ cam_from_arm
        g = np.eye(4)
        cg, sg = math.cos(gamma), math.sin(gamma)
        g[:3, :3] = [[cg, -sg, 0], [sg, cg, 0], [0, 0, 1]]
        return g @ cam_from_arm @ np.linalg.inv(device_from_arm)

    @property
    def camera3d_from_device(self) -> geo.FrameTransform:
        return geo.FrameTransform(self._camera3d_from_device(self.alpha, self.beta, self.gamma, self.isocenter,
                                                             self.source_to_isocenter_vertical_distance,
                                                             self.source_to_isocenter_horizontal_offset, self.rotate_camera_left))

    @property
    def device_from_camera3d(self) -> geo.FrameTransform:
        return self.camera3d_from_device.inv

    @property
    def camera3d_from_world(self) -> geo.FrameTransform:
        return self.camera3d_from_device @ self.device_from_world

    @property
    def isocenter_in_world(self) -> np.ndarray:
        return self.world_from_device @ self.isocenter

    def move_by(self, delta_isocenter=None, delta_alpha: Optional[float] = None, delta_beta: Optional[float] = None, degrees: bool = True):
        if delta_isocenter is not None:
            self.isocenter = self.isocenter + np.asarray(delta_isocenter, dtype=np.float64).reshape(3)
        if delta_alpha is not None:
            self.alpha += _radians(float(delta_alpha), degrees)
        if delta_beta is not None:
            self.beta += _radians(float(delta_beta), degrees)
        self._enforce_bounds()

    def move_to(self, isocenter=None, isocenter_in_world=None, alpha: Optional[float] = None, beta: Optional[float] = None, degrees: bool = True):
        if alpha is not None:
            self.alpha = _radians(float(alpha), degrees)
        if beta is not None:
            self.beta = _radians(float(beta), degrees)
        if isocenter_in_world is not None:
            isocenter = self.device_from_world @ np.asarray(isocenter_in_world, dtype=np.float64).reshape(3)
        if isocenter is not None:
            self.isocenter = np.asarray(isocenter, dtype=np.float64).reshape(3).copy()
        self._enforce_bounds()

    def reposition(self, device_in_world=None):
        self.move_to(isocenter=[0, 0, 0], alpha=0, beta=0, degrees=False)
        if device_in_world is not None:
            self.world_from_device = geo.FrameTransform.from_translation(np.asarray(device_in_world, dtype=np.float64).reshape(3))

    def camera3d_from_world_batch(self, alphas: Sequence[float], betas: Sequence[float], isocenters: Optional[np.ndarray] = None,
                                  degrees: bool = True) -> np.ndarray:
        """``camera3d_from_world`` [n, 4, 4] for n poses at once -- the product ``gamma_rotation @ camera3d_from_arm @ arm_from_device @
        device_from_world`` of the reference (device/mobile_carm.py:223-276) on stacked arrays, no per-view Python loop:
        device_from_arm = [Ry(beta) Rx(alpha) | isocenter] (scipy ``from_euler("xy")``), whose inverse is [R^T | -R^T isocenter]."""
        al = np.asarray([_radians(float(a), degrees) for a in alphas], dtype=np.float64) if not isinstance(alphas, np.ndarray) \
            else (np.deg2rad(alphas.astype(np.float64)) if degrees else alphas.astype(np.float64))
        be = np.asarray([_radians(float(b), degrees) for b in betas], dtype=np.float64) if not isinstance(betas, np.ndarray) \
            else (np.deg2rad(betas.astype(np.float64)) if degrees else betas.astype(np.float64))
        n = al.shape[0]
        iso = np.broadcast_to(self.isocenter, (n, 3)) if isocenters is None else np.asarray(isocenters, dtype=np.float64).reshape(n, 3)
        ca, sa, cb, sb = np.cos(al), np.sin(al), np.cos(be), np.sin(be)
        zero, one = np.zeros(n), np.ones(n)
        rx = np.stack([np.stack([one, zero, zero], -1), np.stack([zero, ca, -sa], -1), np.stack([zero, sa, ca], -1)], -2)
        ry = np.stack([np.stack([cb, zero, sb], -1), np.stack([zero, one, zero], -1), np.stack([-sb, zero, cb], -1)], -2)
        rt = np.transpose(ry @ rx, (0, 2, 1))                                       # R^T
        arm_from_device = np.zeros((n, 4, 4))
        arm_from_device[:, :3, :3] = rt
        arm_from_device[:, :3, 3] = -(rt @ iso[:, :, None])[:, :, 0]
        arm_from_device[:, 3, 3] = 1.0
        cam_from_arm = np.eye(4)
        cam_from_arm[:3, 3] = [0, -self.source_to_isocenter_horizontal_offset, self.source_to_isocenter_vertical_distance]
        if self.rotate_camera_left:
            rz = np.eye(4)
            rz[:3, :3] = [[0, -1, 0], [1, 0, 0], [0, 0, 1]]
            cam_from_arm = rz @ cam_from_arm
        g = np.eye(4)
        cg, sg = math.cos(self.gamma), math.sin(self.gamma)
        g[:3, :3] = [[cg, -sg, 0], [sg, cg, 0], [0, 0, 1]]
        return (g @ cam_from_arm)[None] @ arm_from_device @ self.device_from_world.data[None]

    def camera_projections(self, alphas: Sequence[float], betas: Sequence[float], isocenters: Optional[np.ndarray] = None,
                           degrees: bool = True) -> List[geo.CameraProjection]:
        """A batch of poses without touching the device state (angles are NOT clipped to the device limits).  The matrices come
        from one stacked computation (``camera3d_from_world_batch``); only the thin ``CameraProjection`` wrappers are made per view."""
        k = self.camera_intrinsics
        return [geo.CameraProjection(k, geo.FrameTransform(m)) for m in self.camera3d_from_world_batch(alphas, betas, isocenters, degrees)]

    def __str__(self):
        return f"MobileCArm(isocenter={np.array_str(self.isocenter)}, alpha={math.degrees(self.alpha)}, beta={math.degrees(self.beta)}, degrees=True)"
