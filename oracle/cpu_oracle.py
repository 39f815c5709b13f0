"""TEST INFRASTRUCTURE -- ctypes wrapper of the plain-C oracle (oracle/drr_oracle.c).

The oracle restates the reference's ``projectKernel`` (deepdrr/projector/project_kernel.cu:135-650)
and the host post-processing of ``Projector.project`` (projector.py:691-702, utils/image_utils.py:
18-59) on the CPU.  It is the checker for tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` leg of bench.py -- never a fallback of the product (deepdrr_b200/ does not import
this package).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
MAX_VOL, MAX_MAT = 8, 16


class _Scene(ctypes.Structure):
    _fields_ = [
        ("V", ctypes.c_int), ("M", ctypes.c_int),
        ("density", ctypes.c_void_p * MAX_VOL),
        ("labels", ctypes.c_void_p * MAX_VOL),
        ("shape", (ctypes.c_int * 3) * MAX_VOL),
        ("priority", ctypes.c_int * MAX_VOL),
        ("enabled", ctypes.c_int * MAX_VOL),
        ("W", ctypes.c_int), ("H", ctypes.c_int),
        ("step", ctypes.c_float), ("max_ray_length", ctypes.c_float),
        ("w2i", ctypes.c_float * 9),
        ("src", (ctypes.c_float * 3) * MAX_VOL),
        ("ijk", (ctypes.c_float * 12) * MAX_VOL),
        ("n_bins", ctypes.c_int),
        ("energies", ctypes.c_void_p), ("pdf", ctypes.c_void_p), ("mu", ctypes.c_void_p),
        ("attenuate_outside", ctypes.c_int), ("air_index", ctypes.c_int),
        ("mesh_layers", ctypes.c_int), ("max_hits", ctypes.c_int),
        ("hit_alphas", ctypes.c_void_p), ("hit_facing", ctypes.c_void_p), ("layer_valid", ctypes.c_void_p),
        ("additive", ctypes.c_void_p), ("mesh_mats", ctypes.c_void_p), ("n_mesh_mats", ctypes.c_int),
        ("tex_mode", ctypes.c_int),
    ]


_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/drr_oracle.c with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "drr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
                               "-fno-fast-math", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.drr_oracle_scene_size.restype = ctypes.c_size_t
        assert lib.drr_oracle_scene_size() == ctypes.sizeof(_Scene), "orc_scene layout mismatch"
        lib.drr_oracle_project.argtypes = [ctypes.POINTER(_Scene)] + [ctypes.c_int] * 5 + [ctypes.c_void_p] * 5 + [ctypes.c_int]
        lib.drr_oracle_neglog.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float]
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


class OracleResult:
    def __init__(self, intensity, photon_prob, area, steps, solid):
        self.intensity, self.photon_prob, self.area, self.steps, self.solid = intensity, photon_prob, area, steps, solid


def project(densities: Sequence[np.ndarray], labels_u8: Sequence[np.ndarray], num_materials: int, W: int, H: int,
            step: float, w2i: np.ndarray, src_ijk: np.ndarray, ijk_from_world: np.ndarray, max_ray_length: float,
            energies: np.ndarray, pdf: np.ndarray, mu: np.ndarray, priority: Optional[Sequence[int]] = None,
            enabled: Optional[Sequence[int]] = None, sub: int = 1, u0: int = 0, v0: int = 0,
            Ws: Optional[int] = None, Hs: Optional[int] = None, attenuate_outside: bool = False, air_index: int = 0,
            mesh: Optional[dict] = None, tex_mode: int = 0, want_area: bool = True, want_steps: bool = False,
            want_solid: bool = False, nthreads: int = 0) -> OracleResult:
    """One view on the pixel lattice (u0 + a*sub, v0 + b*sub); images come back as [Hs, Ws]."""
    lib = _load()
    V = len(densities)
    s = _Scene()
    s.V, s.M = V, num_materials
    keep = []
    for v in range(V):
        d = np.ascontiguousarray(densities[v], dtype=np.float32)
        l = np.ascontiguousarray(labels_u8[v], dtype=np.uint8)
        assert d.shape == l.shape
        keep += [d, l]
        s.density[v] = d.ctypes.data
        s.labels[v] = l.ctypes.data
        for a in range(3):
            s.shape[v][a] = d.shape[a]
        s.priority[v] = int(priority[v]) if priority is not None else V - 1 - v
        s.enabled[v] = int(enabled[v]) if enabled is not None else 1
    s.W, s.H, s.step, s.max_ray_length = W, H, float(step), float(max_ray_length)
    w = np.ascontiguousarray(w2i, dtype=np.float32).reshape(9)
    sr = np.ascontiguousarray(src_ijk, dtype=np.float32).reshape(V, 3)
    ij = np.ascontiguousarray(ijk_from_world, dtype=np.float32).reshape(V, 12)
    for k in range(9):
        s.w2i[k] = w[k]
    for v in range(V):
        for k in range(3):
            s.src[v][k] = sr[v, k]
        for k in range(12):
            s.ijk[v][k] = ij[v, k]
    e = np.ascontiguousarray(energies, dtype=np.float32)
    p = np.ascontiguousarray(pdf, dtype=np.float32)
    m = np.ascontiguousarray(mu, dtype=np.float32)
    assert m.size == e.size * num_materials
    s.n_bins, s.energies, s.pdf, s.mu = e.size, e.ctypes.data, p.ctypes.data, m.ctypes.data
    s.attenuate_outside, s.air_index, s.tex_mode = int(attenuate_outside), air_index, tex_mode
    if mesh is not None:
        ha = np.ascontiguousarray(mesh["hit_alphas"], dtype=np.float32)
        hf = np.ascontiguousarray(mesh["hit_facing"], dtype=np.int8)
        lv = np.ascontiguousarray(mesh["layer_valid"], dtype=np.int8)
        keep += [ha, hf, lv]
        s.mesh_layers, s.max_hits = ha.shape[0], ha.shape[2]
        s.hit_alphas, s.hit_facing, s.layer_valid = ha.ctypes.data, hf.ctypes.data, lv.ctypes.data
        if mesh.get("additive") is not None:
            ad = np.ascontiguousarray(mesh["additive"], dtype=np.float32)
            mm = np.ascontiguousarray(mesh["mesh_mats"], dtype=np.int32)
            keep += [ad, mm]
            s.additive, s.mesh_mats, s.n_mesh_mats = ad.ctypes.data, mm.ctypes.data, mm.size
    Ws = Ws if Ws is not None else (W - u0 + sub - 1) // sub
    Hs = Hs if Hs is not None else (H - v0 + sub - 1) // sub
    inten = np.empty((Hs, Ws), dtype=np.float32)
    pp = np.empty((Hs, Ws), dtype=np.float32)
    area = np.empty((num_materials, Hs, Ws), dtype=np.float32) if want_area else None
    steps = np.empty((Hs, Ws), dtype=np.int32) if want_steps else None
    solid = np.empty((Hs, Ws), dtype=np.float32) if want_solid else None
    rc = lib.drr_oracle_project(ctypes.byref(s), u0, v0, sub, Ws, Hs, _ptr(inten), _ptr(pp), _ptr(area), _ptr(steps),
                                _ptr(solid), nthreads)
    if rc != 0:
        raise RuntimeError("oracle: scene exceeds compiled limits")
    return OracleResult(inten, pp, area, steps, solid)


def neglog(images: np.ndarray, epsilon: float = 0.01) -> np.ndarray:
    """``utils.neglog`` (utils/image_utils.py:18-59) per image, float32."""
    lib = _load()
    out = np.array(images, dtype=np.float32, copy=True)
    flat = out.reshape(-1, out.shape[-2] * out.shape[-1]) if out.ndim >= 2 else out.reshape(1, -1)
    for i in range(flat.shape[0]):
        row = np.ascontiguousarray(flat[i])
        lib.drr_oracle_neglog(row.ctypes.data, row.size, epsilon)
        flat[i] = row
    return out
