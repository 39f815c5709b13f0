"""Minimal geometry types at the projector boundary.

The reference takes its geometry from the third-party ``killeengeo`` package (re-exported by
deepdrr/geo/__init__.py:19-66; unpinned in pyproject.toml:40, not vendored, absent here).  The
projection path only touches a handful of attributes (call sites: projector.py:637, 711, 721-722,
803, 813-816, 823): ``CameraProjection.world_from_index``, ``.center_in_world``,
``.intrinsic.sensor_size`` / ``.sensor_width`` / ``.sensor_height`` and
``FrameTransform.inverse()`` / ``.inv`` / ``.toarray()`` / ``@``.  This module provides exactly
those with the published pinhole math

    index_from_world = K [R | t],  world_from_index[:3, :] = R^T K^-1,  center_in_world = -R^T t

so that user code written against ``deepdrr.geo`` keeps working for the projection path.  Anything
that quacks the same way (a real killeengeo object) is accepted by the Projector too.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np


def point(*args) -> np.ndarray:
    a = np.array(args[0] if len(args) == 1 else args, dtype=np.float64).reshape(-1)
    return a


def vector(*args) -> np.ndarray:
    return point(*args)


class FrameTransform:
    """Rigid/affine 3-D transform stored as a 4x4 float64 matrix."""

    def __init__(self, data=None):
        if data is None:
            data = np.eye(4)
        if isinstance(data, FrameTransform):
            data = data.data
        data = np.array(data, dtype=np.float64)
        if data.shape == (3, 4):
            data = np.concatenate([data, [[0, 0, 0, 1]]], axis=0)
        if data.shape != (4, 4):
            raise ValueError(f"FrameTransform needs a 4x4 or 3x4 matrix, got {data.shape}")
        self.data = data

    # constructors -----------------------------------------------------------------------------
    @classmethod
    def identity(cls, dim: int = 3) -> "FrameTransform":
        return cls(np.eye(4))

    @classmethod
    def from_rt(cls, rotation=None, translation=None) -> "FrameTransform":
        m = np.eye(4)
        if rotation is not None:
            m[:3, :3] = np.array(rotation, dtype=np.float64).reshape(3, 3)
        if translation is not None:
            m[:3, 3] = np.array(translation, dtype=np.float64).reshape(-1)[:3]
        return cls(m)

    @classmethod
    def from_scaling(cls, scaling, translation=None) -> "FrameTransform":
        s = np.broadcast_to(np.array(scaling, dtype=np.float64), (3,))
        return cls.from_rt(np.diag(s), translation)

    @classmethod
    def from_translation(cls, translation) -> "FrameTransform":
        return cls.from_rt(None, translation)

    # accessors --------------------------------------------------------------------------------
    @property
    def R(self) -> np.ndarray:
        return self.data[:3, :3]

    @property
    def t(self) -> np.ndarray:
        return self.data[:3, 3]

    def inverse(self) -> "FrameTransform":
        return FrameTransform(np.linalg.inv(self.data))

    @property
    def inv(self) -> "FrameTransform":
        return self.inverse()

    def toarray(self) -> np.ndarray:
        """Top 3x4 block (what projector.py:823-826 flattens into ``ijk_from_world``)."""
        return self.data[:3, :].copy()

    def copy(self) -> "FrameTransform":
        return FrameTransform(self.data.copy())

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    def __matmul__(self, other):
        if isinstance(other, FrameTransform):
            return FrameTransform(self.data @ other.data)
        o = np.asarray(other, dtype=np.float64)
        if o.shape == (3,):  # a point
            return self.data[:3, :3] @ o + self.data[:3, 3]
        if o.shape == (4,):
            return self.data @ o
        if o.shape == (4, 4):
            return FrameTransform(self.data @ o)
        raise TypeError(f"cannot apply FrameTransform to array of shape {o.shape}")

    def transform_vector(self, v) -> np.ndarray:
        return self.data[:3, :3] @ np.asarray(v, dtype=np.float64)

    def __repr__(self):
        return f"FrameTransform(\n{self.data}\n)"


def frame_transform(x=None) -> FrameTransform:
    if x is None:
        return FrameTransform.identity()
    if isinstance(x, FrameTransform):
        return x
    if hasattr(x, "as_matrix"):  # scipy Rotation
        return FrameTransform.from_rt(x.as_matrix())
    x = np.asarray(x, dtype=np.float64)
    if x.shape == (3, 3):
        return FrameTransform.from_rt(x)
    return FrameTransform(x)


class CameraIntrinsicTransform:
    """3x3 pinhole intrinsics + sensor size (W, H)."""

    def __init__(self, data, sensor_height: Optional[int] = None, sensor_width: Optional[int] = None):
        self.data = np.array(data, dtype=np.float64).reshape(3, 3)
        self._sensor_height = sensor_height
        self._sensor_width = sensor_width

    @classmethod
    def from_sizes(cls, sensor_size: Union[int, Tuple[int, int]], pixel_size: Union[float, Tuple[float, float]],
                   source_to_detector_distance: float) -> "CameraIntrinsicTransform":
        """f = SDD / pixel_size, principal point at the sensor centre (W/2, H/2).

        Follows the behaviour the reference relies on at device/mobile_carm.py:142-146 and
        device/simple_device.py:66-78.
        """
        ss = np.broadcast_to(np.array(sensor_size), (2,))
        ps = np.broadcast_to(np.array(pixel_size, dtype=np.float64), (2,))
        fx, fy = source_to_detector_distance / ps[0], source_to_detector_distance / ps[1]
        k = np.array([[fx, 0, ss[0] / 2], [0, fy, ss[1] / 2], [0, 0, 1]], dtype=np.float64)
        return cls(k, sensor_height=int(ss[1]), sensor_width=int(ss[0]))

    @property
    def fx(self): return float(self.data[0, 0])
    @property
    def fy(self): return float(self.data[1, 1])
    @property
    def cx(self): return float(self.data[0, 2])
    @property
    def cy(self): return float(self.data[1, 2])

    @property
    def sensor_width(self) -> int:
        return int(self._sensor_width if self._sensor_width is not None else np.ceil(2 * self.cx))

    @property
    def sensor_height(self) -> int:
        return int(self._sensor_height if self._sensor_height is not None else np.ceil(2 * self.cy))

    @property
    def sensor_size(self) -> Tuple[int, int]:
        return (self.sensor_width, self.sensor_height)

    @property
    def inv(self) -> np.ndarray:
        return np.linalg.inv(self.data)


class CameraProjection:
    """intrinsic (index_from_camera2d) + extrinsic (camera3d_from_world)."""

    def __init__(self, intrinsic: CameraIntrinsicTransform, extrinsic: FrameTransform):
        self.index_from_camera2d = intrinsic if isinstance(intrinsic, CameraIntrinsicTransform) \
            else CameraIntrinsicTransform(intrinsic)
        self.camera3d_from_world = frame_transform(extrinsic)

    @property
    def intrinsic(self) -> CameraIntrinsicTransform:
        return self.index_from_camera2d

    @property
    def extrinsic(self) -> FrameTransform:
        return self.camera3d_from_world

    @property
    def sensor_width(self) -> int:
        return self.intrinsic.sensor_width

    @property
    def sensor_height(self) -> int:
        return self.intrinsic.sensor_height

    @property
    def index_from_world(self) -> np.ndarray:
        return self.intrinsic.data @ self.camera3d_from_world.data[:3, :]

    @property
    def world_from_index(self) -> np.ndarray:
        """4x3: maps homogeneous pixel (u, v, 1) to the world-space ray vector (w = 0 row last)."""
        r = self.camera3d_from_world.data[:3, :3]
        m = r.T @ np.linalg.inv(self.intrinsic.data)
        return np.concatenate([m, np.zeros((1, 3))], axis=0)

    @property
    def center_in_world(self) -> np.ndarray:
        e = self.camera3d_from_world.data
        return -e[:3, :3].T @ e[:3, 3]


def pose_arrays_batch(projs: Sequence, volumes: Sequence) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per-view kernel inputs for a batch of views, exactly as the reference derives them one view at a time
    (projector.py:802-831): ``world_from_index`` (n, 9) f32, ``source_ijk`` (n, V, 3) f32, ``ijk_from_world`` (n, V, 12) f32.

    All views go through stacked float64 NumPy operations (no per-view Python loop for this module's own
    ``CameraProjection``); foreign projection objects (killeengeo) are read through their ``world_from_index`` /
    ``center_in_world`` properties.  ``pose_arrays`` is this function with n = 1, so both give identical bits.
    """
    n, V = len(projs), len(volumes)
    if n and all(type(p) is CameraProjection for p in projs):
        e = np.stack([p.camera3d_from_world.data for p in projs])                 # (n, 4, 4)
        k = np.stack([p.index_from_camera2d.data for p in projs])                 # (n, 3, 3)
        rt = np.transpose(e[:, :3, :3], (0, 2, 1))
        w64 = rt @ np.linalg.inv(k)                                               # world_from_index[:3] = R^T K^-1
        c = -(rt @ e[:, :3, 3:4])[:, :, 0]                                        # center_in_world = -R^T t
    else:
        w64 = np.stack([np.asarray(p.world_from_index, dtype=np.float64)[:-1, :] for p in projs]) if n else np.zeros((0, 3, 3))
        c = np.stack([np.asarray(p.center_in_world, dtype=np.float64).reshape(-1)[:3] for p in projs]) if n else np.zeros((0, 3))
    w2i = np.ascontiguousarray(w64.astype(np.float32).reshape(n, 9))
    src = np.zeros((n, V, 3), dtype=np.float32)
    a = np.zeros((n, V, 12), dtype=np.float32)
    for i, v in enumerate(volumes):
        t = v.IJK_from_world
        m = np.asarray(t.toarray() if hasattr(t, "toarray") else t, dtype=np.float64)[:3, :]
        src[:, i, :] = (c @ m[:, :3].T + m[:, 3]).astype(np.float32)
        a[:, i, :] = m.astype(np.float32).reshape(12)
    return w2i, src, a


def pose_arrays(proj, volumes: Sequence) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One view: ``world_from_index`` (9,) f32, ``source_ijk`` (V, 3) f32, ``ijk_from_world`` (V, 12) f32."""
    w2i, src, a = pose_arrays_batch([proj], volumes)
    return w2i[0], src[0], a[0]
