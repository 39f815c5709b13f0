"""TEST INFRASTRUCTURE -- ctypes driver for the reference's own GPU kernel (oracle/_ref).

``oracle/_ref/*.cubin`` are built by ``oracle/Makefile`` from the reference's unmodified
``deepdrr/projector/project_kernel.cu``; ``libref_harness.so`` (oracle/ref_harness.cu) re-creates the
texture setup and the 37-argument launch of ``deepdrr/projector/projector.py:116-257, 718-774``.
This is "the reference's projector on the same B200": the primary parity checker for the CUDA path
and the reference arm of bench.py.  It must never be imported by the product package.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_lib = None


def available() -> bool:
    return os.path.exists(os.path.join(_REF, "libref_harness.so"))


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(os.path.join(_REF, "libref_harness.so"))
        lib.ref_last_error.restype = ctypes.c_char_p
        lib.ref_create.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        lib.ref_add_volume.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float]
        lib.ref_set_spectrum.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.ref_project.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_float)]
        lib.ref_destroy.argtypes = [ctypes.c_void_p]
        lib.ref_want_solid.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.ref_fetch_solid.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.ref_set_mesh.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_int]
        lib.ref_tide.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float]
        _lib = lib
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError("ref harness: " + _load().ref_last_error().decode())


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class RefProjector:
    """One compiled (NUM_VOLUMES, NUM_MATERIALS) instance of the reference kernel."""

    def __init__(self, densities: Sequence[np.ndarray], labels_u8: Sequence[np.ndarray], num_materials: int,
                 spacings: Optional[Sequence[Sequence[float]]] = None, device: int = 0, lineint: bool = False, variant: str = ""):
        """variant: "" (plain), "mesh" (MESH_ADDITIVE_ENABLED=1) or "att0" (ATTENUATE_OUTSIDE_VOLUME=1, AIR_INDEX=0)."""
        lib = _load()
        V = len(densities)
        kind = ("lineint" if lineint else "project") + (("_" + variant) if variant else "")
        path = os.path.join(_REF, f"ref_{kind}_V{V}_M{num_materials}.cubin")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: add the (V, M) pair to oracle/Makefile CONFIGS and rebuild")
        self.V, self.M, self.lineint = V, num_materials, lineint
        self.h = ctypes.c_void_p()
        _chk(lib.ref_create(path.encode(), device, V, num_materials, 32, 2, ctypes.byref(self.h)))
        for v in range(V):
            d = np.ascontiguousarray(densities[v], dtype=np.float32)
            l = np.ascontiguousarray(labels_u8[v], dtype=np.uint8)
            sp = (1.0, 1.0, 1.0) if spacings is None else spacings[v]
            _chk(lib.ref_add_volume(self.h, _p(d), _p(l), d.shape[0], d.shape[1], d.shape[2], sp[0], sp[1], sp[2]))
        self._spectrum = None

    def set_mesh(self, mesh: dict, npix: int):
        """Feed the mesh buffers projectKernel consumes (same dict as oracle.cpu_oracle.project(mesh=...))."""
        ha = np.ascontiguousarray(mesh["hit_alphas"], dtype=np.float32)
        hf = np.ascontiguousarray(mesh["hit_facing"], dtype=np.int8)
        lv = np.ascontiguousarray(mesh["layer_valid"], dtype=np.int8)
        ad = np.ascontiguousarray(mesh["additive"], dtype=np.float32)
        mm = np.ascontiguousarray(mesh["mesh_mats"], dtype=np.int32)
        assert ha.shape[0] == 2 and ha.shape[2] == 32, "the cubins are compiled with MESH_LAYERS=2, MAX_MESH_HITS=32"
        _chk(_load().ref_set_mesh(self.h, _p(ha), _p(hf), _p(lv), _p(ad), _p(mm), mm.size, npix))

    def set_spectrum(self, energies: np.ndarray, pdf: np.ndarray, mu: np.ndarray):
        e = np.ascontiguousarray(energies, dtype=np.float32)
        p = np.ascontiguousarray(pdf, dtype=np.float32)
        m = np.ascontiguousarray(mu, dtype=np.float32)
        assert m.size == e.size * self.M
        _chk(_load().ref_set_spectrum(self.h, e.size, _p(e), _p(p), _p(m)))
        self._spectrum = (e, p, m)

    def project(self, W: int, H: int, step: float, w2i: np.ndarray, src_ijk: np.ndarray, ijk_from_world: np.ndarray,
                max_ray_length: float, priority: Optional[Sequence[int]] = None,
                enabled: Optional[Sequence[int]] = None, threads: int = 8, transpose: bool = True,
                fetch: bool = True) -> Tuple[Optional[np.ndarray], Optional[np.ndarray], float]:
        """Returns (intensity [H, W], photon_prob [H, W], kernel_ms) for one view."""
        pr = np.ascontiguousarray(priority if priority is not None else [self.V - 1 - i for i in range(self.V)], dtype=np.int32)
        en = np.ascontiguousarray(enabled if enabled is not None else [1] * self.V, dtype=np.int32)
        w = np.ascontiguousarray(w2i, dtype=np.float32).reshape(9)
        s = np.ascontiguousarray(src_ijk, dtype=np.float32).reshape(self.V * 3)
        a = np.ascontiguousarray(ijk_from_world, dtype=np.float32).reshape(self.V * 12)
        inten = np.empty((H, W) if transpose else (W, H), dtype=np.float32) if fetch else None
        pprob = np.empty_like(inten) if fetch else None
        ms = ctypes.c_float(0)
        _chk(_load().ref_project(self.h, W, H, float(step), _p(pr), _p(en), _p(s), float(max_ray_length), _p(w), _p(a),
                                 threads, int(transpose), _p(inten), _p(pprob), ctypes.byref(ms)))
        return inten, pprob, ms.value

    def solid_angle(self, W: int, H: int, w2i: np.ndarray, src_ijk: np.ndarray, ijk_from_world: np.ndarray, max_ray_length: float) -> np.ndarray:
        """[H, W] solid angle per pixel as the reference kernel's ``calculate_solid_angle`` writes it
        (project_kernel.cu:14-133, 213-216; a function of world_from_index and the pixel only)."""
        assert self._spectrum is not None, "set_spectrum first (the kernel runs as a whole)"
        _chk(_load().ref_want_solid(self.h, 1))
        try:
            self.project(W, H, 1.0e6, w2i, src_ijk, ijk_from_world, max_ray_length, fetch=False)  # huge step: the march is over at once
            out = np.empty((H, W), dtype=np.float32)
            _chk(_load().ref_fetch_solid(self.h, _p(out), 1))
        finally:
            _load().ref_want_solid(self.h, 0)
        return out

    def line_integrals(self, W, H, step, w2i, src_ijk, ijk_from_world, max_ray_length, priority=None, enabled=None):
        """[M, H, W] per-material area densities (g/cm^2) straight from the reference's march."""
        assert self.lineint, "construct with lineint=True"
        out = np.empty((self.M, H, W), dtype=np.float32)
        for m in range(self.M):
            mu = np.zeros(self.M, dtype=np.float32)
            mu[m] = 1.0
            self.set_spectrum(np.ones(1, np.float32), -np.ones(1, np.float32), mu)
            _, pp, _ = self.project(W, H, step, w2i, src_ijk, ijk_from_world, max_ray_length, priority, enabled)
            out[m] = pp
        return out

    def close(self):
        if self.h:
            _load().ref_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shutdown():
    """Wait for the device to go idle (every RefProjector should have been closed)."""
    if _lib is not None:
        _lib.ref_shutdown()


def ref_tide(ts_peel: np.ndarray, far_limit: float):
    """The reference's kernelTide on ``ts_peel`` [n_rays, 32] (peel layout); returns (ts, facing) cleaned."""
    ts = np.ascontiguousarray(ts_peel, dtype=np.float32).copy()
    assert ts.ndim == 2 and ts.shape[1] == 32
    facing = np.zeros(ts.shape, dtype=np.int8)
    _chk(_load().ref_tide(os.path.join(_REF, "ref_peel.cubin").encode(), _p(ts), _p(facing), ts.shape[0], float(far_limit)))
    return ts, facing
