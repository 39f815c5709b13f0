"""Volume data contract at the projector boundary (host side, NumPy).

Only what the projection path consumes from the reference's ``deepdrr.vol`` package:

* ``Volume.data``       float32 ``[Ni, Nj, Nk]`` densities (g/cm^3)          (vol/volume.py:178-179)
* ``Volume.materials``  ``(dict name -> id, uint16 labels [Ni, Nj, Nk])``    (vol/volume.py:181-184)
* pose: ``world_from_IJK = world_from_anatomical @ anatomical_from_IJK``      (vol/renderable.py:58-84)
* ``enabled`` flag                                                           (vol/volume.py:190-192)

plus the two conversions the synthetic configs need: HU -> density (vol/volume.py:338-351) and
threshold segmentation (load_dicom.py:132-143), and the NIfTI / NRRD loaders (``Volume.from_nifti`` /
``from_nrrd``, vol/volume.py:581-696, 848-895; file parsing in ``formats.py``).  DICOM, mesh tooling and the
V-Net segmentation are out of scope.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple, Union

import numpy as np

from . import geo


class Renderable:
    """Pose bookkeeping shared by volumes and meshes (reference: vol/renderable.py:33-84)."""

    def __init__(self, anatomical_from_IJK=None, world_from_anatomical=None, anatomical_from_ijk=None,
                 enabled: bool = True):
        if anatomical_from_ijk is not None:
            anatomical_from_IJK = anatomical_from_ijk
        self.anatomical_from_IJK = geo.frame_transform(anatomical_from_IJK)
        self.world_from_anatomical = geo.frame_transform(world_from_anatomical)
        self.enabled = enabled

    def set_enabled(self, enabled: bool) -> None:
        self.enabled = enabled

    @property
    def anatomical_from_ijk(self):
        return self.anatomical_from_IJK

    @property
    def world_from_IJK(self) -> geo.FrameTransform:
        return self.world_from_anatomical @ self.anatomical_from_IJK

    @property
    def world_from_ijk(self) -> geo.FrameTransform:
        return self.world_from_IJK

    @property
    def IJK_from_world(self) -> geo.FrameTransform:
        return self.world_from_IJK.inverse()

    @property
    def ijk_from_world(self) -> geo.FrameTransform:
        return self.world_from_IJK.inv

    # pose helpers used by user scripts (subset of vol/renderable.py:150-234)
    def place_center(self, x) -> None:
        x = np.asarray(x, dtype=np.float64).reshape(-1)[:3]
        center_anat = self.anatomical_from_IJK @ (np.array(self.shape, dtype=np.float64) / 2)
        center_world = self.world_from_anatomical @ center_anat
        self.translate(x - center_world)

    def translate(self, t) -> None:
        t = np.asarray(t, dtype=np.float64).reshape(-1)[:3]
        self.world_from_anatomical = geo.FrameTransform.from_translation(t) @ self.world_from_anatomical

    def rotate(self, rotation, center=None) -> None:
        r = rotation.as_matrix() if hasattr(rotation, "as_matrix") else np.asarray(rotation, dtype=np.float64)
        c = np.zeros(3) if center is None else np.asarray(center, dtype=np.float64).reshape(-1)[:3]
        m = geo.FrameTransform.from_translation(c) @ geo.FrameTransform.from_rt(r) @ geo.FrameTransform.from_translation(-c)
        self.world_from_anatomical = m @ self.world_from_anatomical


def convert_hounsfield_to_density(hu_values: np.ndarray, smooth_air: bool = False) -> np.ndarray:
    """Two-segment linear HU -> g/cm^3 map, clamped at 0 (reference: vol/volume.py:338-351)."""
    if smooth_air:
        hu_values[hu_values <= -900] = -1000
    return np.maximum(np.minimum(0.001029 * hu_values + 1.030, 0.0005886 * hu_values + 1.03), 0)


def segment_materials_thresholding(hu_values: np.ndarray) -> Dict[str, np.ndarray]:
    """air <= -800 < soft tissue <= 350 < bone (reference: load_dicom.py:132-143)."""
    materials = {}
    materials["air"] = hu_values <= -800
    materials["soft tissue"] = (-800 < hu_values) * (hu_values <= 350)
    materials["bone"] = 350 < hu_values
    return materials


def format_materials(materials: Dict[str, np.ndarray]) -> Tuple[Dict[str, int], np.ndarray]:
    """dict of masks -> (name -> id, uint16 labels); later masks overwrite earlier ones and
    unlabeled voxels stay 0 (reference: vol/volume.py:955-992)."""
    combined = None
    mdict: Dict[str, int] = {}
    for mat_id, mat in enumerate(materials):
        if combined is None:
            combined = np.zeros(materials[mat].shape, dtype=np.uint16)
        combined[materials[mat] > 0] = mat_id
        mdict[mat] = mat_id
    return mdict, combined


class Volume(Renderable):
    """Density volume + material labels + pose (reference: vol/volume.py:142-192)."""

    def __init__(self, data: np.ndarray,
                 materials: Union[Dict[str, np.ndarray], Tuple[Dict[str, int], np.ndarray]],
                 anatomical_from_IJK=None, world_from_anatomical=None,
                 anatomical_coordinate_system: Optional[str] = None, enabled: bool = True, **kwargs):
        Renderable.__init__(self, anatomical_from_IJK, world_from_anatomical, enabled=enabled, **kwargs)
        assert np.ndim(data) == 3, "Volume data must be 3D."
        self.data = np.array(data).astype(np.float32)
        if isinstance(materials, tuple):
            self.materials = materials[0], np.asarray(materials[1]).astype(np.uint16)
        else:
            self.materials = format_materials(materials)
        assert anatomical_coordinate_system in ["LPS", "RAS", None]
        self.anatomical_coordinate_system = anatomical_coordinate_system

    @classmethod
    def from_hu(cls, hu_values: np.ndarray, anatomical_from_IJK=None, world_from_anatomical=None, **kwargs):
        """HU volume -> Volume with threshold segmentation (reference: vol/volume.py:194-260 with
        ``use_thresholding=True``)."""
        data = convert_hounsfield_to_density(hu_values)
        materials = segment_materials_thresholding(hu_values)
        return cls(data, materials, anatomical_from_IJK, world_from_anatomical, **kwargs)

    @classmethod
    def from_nifti(cls, path, world_from_anatomical=None, use_thresholding: bool = True, use_cached: bool = True,
                   save_cache: bool = False, cache_dir=None, materials=None, segmentation: bool = False, label=None,
                   binarize: bool = False, density_kwargs: Optional[dict] = None, **kwargs):
        """Load a CT (HU) or a segmentation from a NIfTI-1 file (reference: vol/volume.py:581-696).

        ``anatomical_from_IJK`` is the file's affine (RAS).  Only threshold segmentation exists here
        (``use_thresholding=False`` is the reference's V-Net, out of scope); the cache arguments are accepted and unused.
        ``materials`` may map names to boolean arrays or to NIfTI paths of masks (> 0).
        """
        from . import formats
        values, affine, _ = formats.read_nifti(path)
        anatomical_from_IJK = geo.FrameTransform(affine)
        if segmentation:
            if label is None:
                seg = values > 0
            elif isinstance(label, (int, np.integer, float)):
                seg = values == label
            elif isinstance(label, list):
                seg = np.isin(values, label)
            else:
                raise ValueError(f"Invalid label: {label}")
            materials = dict(bone=seg)
            data = seg.astype(np.float32) if binarize else values.astype(np.float32)
        else:
            data = convert_hounsfield_to_density(values, **(density_kwargs or {}))
            if materials is None:
                if not use_thresholding:
                    raise NotImplementedError("only threshold segmentation is available (the V-Net segmenter is out of scope)")
                materials = segment_materials_thresholding(values)
            else:
                materials = dict(materials)
                for m in materials:
                    if isinstance(materials[m], (str, os.PathLike)):
                        if not os.path.exists(materials[m]):
                            raise ValueError(f"Could not find material {m} at {materials[m]}")
                        materials[m] = formats.read_nifti(materials[m])[0] > 0
                    else:
                        materials[m] = np.asarray(materials[m]).astype(bool)
        return cls(data, materials, anatomical_from_IJK=anatomical_from_IJK, world_from_anatomical=world_from_anatomical,
                   anatomical_coordinate_system="RAS", **kwargs)

    @classmethod
    def from_nrrd(cls, path, world_from_anatomical=None, use_thresholding: bool = True, use_cached: bool = True, cache_dir=None, **kwargs):
        """Load a CT (HU) from an NRRD file (reference: vol/volume.py:848-895).

        As in the reference, the 3x4 ``[space directions | space origin]`` is used as it stands -- axis i's direction
        ends up in ROW i -- which equals the usual column convention only for axis-aligned (or symmetric) directions.
        """
        from . import formats
        values, header = formats.read_nrrd(path)
        if values.ndim != 3:
            raise ValueError(f"{path}: expected a 3-D volume, got shape {values.shape}")
        m = np.concatenate([np.asarray(header["space directions"], dtype=np.float64), np.asarray(header["space origin"], dtype=np.float64).reshape(-1, 1)], axis=1)
        anatomical_from_ijk = geo.FrameTransform(np.concatenate([m, [[0, 0, 0, 1]]], axis=0))
        if not use_thresholding:
            raise NotImplementedError("only threshold segmentation is available (the V-Net segmenter is out of scope)")
        data = convert_hounsfield_to_density(values)
        materials = segment_materials_thresholding(values)
        system = {"right-anterior-superior": "RAS", "left-posterior-superior": "LPS"}.get(header.get("space", "right-anterior-superior"))
        return cls(data, materials, anatomical_from_ijk, world_from_anatomical, anatomical_coordinate_system=system, **kwargs)

    @property
    def shape(self) -> Tuple[int, int, int]:
        return self.data.shape

    @property
    def spacing(self) -> np.ndarray:
        return np.abs(np.array(self.anatomical_from_IJK.R)).max(axis=0)

    def __array__(self, dtype=None, copy=None) -> np.ndarray:
        return self.data if dtype is None else self.data.astype(dtype)

    def get_center(self) -> np.ndarray:
        return self.anatomical_from_IJK @ (np.array(self.shape, dtype=np.float64) / 2)


class HUVolume(Volume):
    """A CT given in Hounsfield units whose density / threshold segmentation are computed on the GPU at
    ``Projector.initialize``; the host arrays are only built if somebody asks for them.

    Equivalent to ``Volume.from_hu(hu, ...)`` -- same materials (air, soft tissue, bone), same float32 arithmetic.
    """

    def __init__(self, hu_values: np.ndarray, anatomical_from_IJK=None, world_from_anatomical=None, enabled: bool = True, **kwargs):
        Renderable.__init__(self, anatomical_from_IJK, world_from_anatomical, enabled=enabled, **kwargs)
        assert np.ndim(hu_values) == 3, "Volume data must be 3D."
        self.hu = np.ascontiguousarray(hu_values, dtype=np.float32)
        self.anatomical_coordinate_system = None
        self._host = None

    def _materialise(self):
        if self._host is None:
            self._host = (convert_hounsfield_to_density(self.hu).astype(np.float32), format_materials(segment_materials_thresholding(self.hu)))
        return self._host

    @property
    def data(self):
        return self._materialise()[0]

    @property
    def materials(self):
        if self._host is None:  # the name -> id map is known without touching the voxels
            return {"air": 0, "soft tissue": 1, "bone": 2}, _LazyLabels(self)
        return self._host[1]

    @property
    def shape(self):
        return self.hu.shape


class _LazyLabels:
    """Stands in for the uint16 label array of an HUVolume until it is really needed."""

    def __init__(self, owner):
        self._owner = owner

    def __array__(self, dtype=None, copy=None):
        a = self._owner._materialise()[1][1]
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, k):
        return self._owner._materialise()[1][1][k]


# Default mesh densities in g/cm^3 by material name (reference: pyrenderdrr/material.py:5-20).
DEFAULT_MESH_DENSITIES = {
    "bone": 1.92, "soft tissue": 1.0, "tissue_soft": 1.0, "blood": 1.06, "muscle": 1.06, "air": 0.0012, "iron": 7.87,
    "lead": 11.34, "copper": 8.96, "lung": 0.26, "titanium": 4.51, "teflon": 2.2, "polyethylene": 0.94, "concrete": 2.3,
}


class Mesh(Renderable):
    """Triangle mesh with a DRR material (reference: vol/mesh.py:40-183 + pyrenderdrr/material.py:22-124).

    One ``Mesh`` is one primitive: ``vertices`` [n, 3] in mesh-local (IJK) coordinates, ``faces`` [m, 3] with
    outward normals counter-clockwise, material name (must be a known ``Material``), ``density`` in g/cm^3
    (default by name), ``additive`` / ``subtractive`` flags and ``layer`` (pyrenderdrr/material.py:38-44).
    Meshes must be watertight and the source must be outside them (README.md:257).
    """

    def __init__(self, vertices, faces, material: str = "iron", density: Optional[float] = None, additive: bool = True,
                 subtractive: bool = False, layer: int = 0, tag: Optional[str] = None, anatomical_from_IJK=None,
                 world_from_anatomical=None, enabled: bool = True):
        Renderable.__init__(self, anatomical_from_IJK, world_from_anatomical, enabled=enabled)
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        self.faces = np.ascontiguousarray(faces, dtype=np.int64).reshape(-1, 3)
        if self.faces.size and (self.faces.min() < 0 or self.faces.max() >= len(self.vertices)):
            raise ValueError("face index out of range")
        self.material = material
        if density is None:
            if material not in DEFAULT_MESH_DENSITIES:
                raise ValueError(f"no default density for material {material!r}; pass density=")
            density = DEFAULT_MESH_DENSITIES[material]
        self.density = float(density)
        self.additive, self.subtractive, self.layer, self.tag = bool(additive), bool(subtractive), int(layer), tag

    @property
    def triangles(self) -> np.ndarray:
        """[m, 3, 3] float32 vertex coordinates per triangle (mesh-local)."""
        return self.vertices[self.faces]

    @property
    def shape(self):
        return tuple(np.ptp(self.vertices, axis=0)) if len(self.vertices) else (0, 0, 0)

    def get_center(self) -> np.ndarray:
        return self.anatomical_from_IJK @ self.vertices.mean(axis=0).astype(np.float64)

    @property
    def get_bounding_AABB(self):
        """(min corner, max corner) of the vertices, mesh-local (reference vol/mesh.py:142-147)."""
        return self.vertices.min(axis=0).astype(np.float64), self.vertices.max(axis=0).astype(np.float64)

    @property
    def get_loose_bounding_sphere(self):
        """(centre of the AABB, distance to the farthest vertex), mesh-local (reference vol/mesh.py:149-162)."""
        lo, hi = self.get_bounding_AABB
        center = (lo + hi) / 2
        return center, float(np.max(np.linalg.norm(self.vertices.astype(np.float64) - center, axis=1)))

    @classmethod
    def from_stl(cls, path, **kwargs) -> "Mesh":
        """Binary or ASCII STL reader (the reference loads STLs through trimesh / pyvista, vol/mesh.py:91-140)."""
        raw = open(path, "rb").read()
        n = int.from_bytes(raw[80:84], "little") if len(raw) >= 84 else -1
        if n >= 0 and len(raw) == 84 + 50 * n:
            rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
            tris = rec["v"].astype(np.float32)
        else:
            import re
            nums = re.findall(rb"vertex\s+(\S+)\s+(\S+)\s+(\S+)", raw)
            tris = np.array(nums, dtype=np.float32).reshape(-1, 3, 3)
        verts = tris.reshape(-1, 3)
        faces = np.arange(len(verts)).reshape(-1, 3)
        return cls(verts, faces, **kwargs)
