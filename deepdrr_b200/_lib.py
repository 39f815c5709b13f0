"""ctypes binding of libdrr_b200.so (include/drr_b200.h).

The library is built in-tree (``deepdrr_b200/libdrr_b200.so``, see ``__graft_entry__.build`` or
``make -C deepdrr_b200/csrc``).  There is no CPU fallback: if the library is missing, or no CUDA
device is visible, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DRR_B200_LIB") or os.path.join(_HERE, "libdrr_b200.so")  # env override: A/B builds during tuning

OK, E_INVALID, E_CUDA, E_STATE, E_NOMEM = 0, -1, -2, -3, -4
MEM_HOST, MEM_DEVICE = 0, 1
MESH_QUERY_HITS, MESH_QUERY_TRAVEL, MESH_QUERY_SEG = 0, 1, 2
SAMPLER_ALU, SAMPLER_TEX, SAMPLER_HYBRID = 0, 1, 2
POST_NEGLOG, POST_NOISE, POST_CLIP, POST_COLLECTED = 1, 2, 4, 8
TUNE_TEX_EIGHTHS, TUNE_KERNEL_VARIANT, TUNE_PIPELINE, TUNE_LANE_QUADS, TUNE_RAYS_PER_LANE = 0, 1, 2, 3, 4
MAX_VOLUMES, MAX_MATERIALS = 8, 16

# every symbol include/drr_b200.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "drr_create", "drr_destroy", "drr_last_error", "drr_set_stream", "drr_set_spectrum", "drr_add_volume", "drr_add_volume_hu",
    "drr_clear_volumes", "drr_set_priorities", "drr_set_march", "drr_set_tuning", "drr_set_mesh_buffers", "drr_set_meshes", "drr_set_mesh_poses", "drr_mesh_clean_hits", "drr_mesh_query", "drr_set_scatter_tables", "drr_scatter", "drr_postprocess",
    "drr_project", "drr_last_timing", "drr_last_sample_count", "drr_last_window_samples", "drr_host_alloc", "drr_host_free", "drr_launch_count", "drr_synchronize", "drr_version",
]

_lib = None


class DrrLibraryMissing(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DrrLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(deepdrr_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cf, cu = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint
    lib.drr_version.restype = ctypes.c_char_p
    lib.drr_last_error.restype = ctypes.c_char_p
    lib.drr_last_error.argtypes = [vp]
    lib.drr_create.argtypes = [ci, ctypes.POINTER(vp)]
    lib.drr_destroy.argtypes = [vp]
    lib.drr_set_stream.argtypes = [vp, vp]
    lib.drr_set_spectrum.argtypes = [vp, ci, ci, vp, vp, vp]
    lib.drr_add_volume.argtypes = [vp, vp, vp, ci, ci, ci, ci, cu, ctypes.POINTER(ci)]
    lib.drr_add_volume_hu.argtypes = [vp, vp, ci, ci, ci, ci, vp, cu, ctypes.POINTER(ci)]
    lib.drr_clear_volumes.argtypes = [vp]
    lib.drr_set_priorities.argtypes = [vp, vp, vp, ci]
    lib.drr_set_march.argtypes = [vp, cf, ci, ci, ci]
    lib.drr_set_tuning.argtypes = [vp, ci, ci]
    lib.drr_set_mesh_buffers.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, ci, ci]
    lib.drr_set_meshes.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, ci, ci]
    lib.drr_set_mesh_poses.argtypes = [vp, ci, vp, vp, cf]
    lib.drr_mesh_clean_hits.argtypes = [vp, vp, vp, ci, ci, cf, ci]
    lib.drr_mesh_query.argtypes = [vp, ci, ci, ci, vp, vp, vp, ci]
    lib.drr_set_scatter_tables.argtypes = [vp, ci, ci, cf, cf, vp, vp, vp, vp, vp, vp, vp]
    lib.drr_scatter.argtypes = [vp, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_uint64, ci, ci, vp, vp, vp, vp, vp, vp, ci]
    lib.drr_postprocess.argtypes = [vp, vp, vp, ci, ci, ci, cu, cf, cf, ctypes.c_uint64, ci]
    lib.drr_project.argtypes = [vp, ci, ci, ci, vp, vp, vp, cf, cu, cf, cf, cf, ctypes.c_uint64, vp, vp, vp, ci]
    lib.drr_last_timing.argtypes = [vp, vp]
    lib.drr_last_sample_count.argtypes = [vp, ctypes.POINTER(ctypes.c_ulonglong)]
    lib.drr_last_window_samples.argtypes = [vp, ctypes.POINTER(ctypes.c_ulonglong)]
    lib.drr_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.drr_host_free.argtypes = [vp]
    lib.drr_launch_count.argtypes = [vp, ctypes.POINTER(ctypes.c_ulonglong)]
    lib.drr_synchronize.argtypes = [vp]
    _lib = lib
    return lib


_EXC = {E_INVALID: ValueError, E_CUDA: RuntimeError, E_STATE: RuntimeError, E_NOMEM: MemoryError}


def check(rc: int, handle: Optional[ctypes.c_void_p] = None):
    """Map a DRR_E_* code to the exception type the reference raises for the same condition."""
    if rc == OK:
        return
    msg = load().drr_last_error(handle).decode(errors="replace")
    raise _EXC.get(rc, RuntimeError)(f"libdrr_b200: {msg} (code {rc})")


def ptr(a) -> Optional[int]:
    """Address of a NumPy array / torch tensor / raw int; None passes NULL."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(f"cannot take the address of {type(a)}")
