"""Monte Carlo scatter for the Projector (north_star kernel 3), host side.

The reference rejects ``scatter_num > 0`` (projector.py:530-531) because its transport kernel was removed; what
it still ships are the MC-GPU data tables (mcgpu_mfp_data.py, mcgpu_rita_samplers.py, mcgpu_compton_data.py,
mcgpu_density.py) and the material-name mapping of conv_to_mcgpu.py:15-35.  Those tables are packed in
``data/mcgpu_tables.npz`` (tools/gen_scatter_tables.py); the transport runs in libdrr_b200 (csrc/drr_scatter.cu).

Photons shard over GPUs: rank r of w simulates the photon ids ``[r*n/w, (r+1)*n/w)`` of the same Philox stream and
the integer tallies are summed with one NCCL all-reduce -- the sum is exact, so the result does not depend on w.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np

from . import _lib

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "mcgpu_tables.npz")

# deepdrr material name -> MC-GPU material (conv_to_mcgpu.py:15-35; the extra identities are table names)
MCGPU_NAME = {"air": "air", "bone": "bone", "soft tissue": "soft tissue", "lung": "lung", "titanium": "titanium", "muscle": "muscle",
              "blood": "blood", "water": "water", "adipose": "adipose", "PMMA": "PMMA"}

_tables = None


def load_tables():
    global _tables
    if _tables is None:
        z = np.load(_DATA)
        _tables = {k: z[k] for k in z.files}
    return _tables


def setup(projector) -> None:
    """Upload the interaction tables for ``projector``'s materials (any number of volumes)."""
    if len(projector.volumes) < 1:
        raise ValueError("scatter simulation needs at least one volume")
    t = load_tables()
    names = [str(n) for n in t["names"]]
    mat_of_label = []
    for m in projector.all_materials:
        if m not in MCGPU_NAME or MCGPU_NAME[m] not in names:
            raise ValueError(f"UNSUPPORTED MATERIAL FOR MCGPU: {m}")  # conv_to_mcgpu.py:24-33
        mat_of_label.append(names.index(MCGPU_NAME[m]))
    from .scene import remap_labels

    rho_max = np.zeros(len(projector.all_materials), dtype=np.float32)   # largest density per material over all volumes: the majorant
    for vol in projector.volumes:
        labels = remap_labels(vol, projector.all_materials)
        dens = np.asarray(vol.data)
        for l in range(len(projector.all_materials)):
            sel = labels == l
            if np.any(sel):
                rho_max[l] = max(rho_max[l], np.float32(dens[sel].max()))
    e = t["energy_eV"].astype(np.float64)
    mfp = np.ascontiguousarray(t["mfp_mm"], dtype=np.float32)
    rita = np.ascontiguousarray(t["rita"], dtype=np.float32)
    comp = np.ascontiguousarray(t["compton"], dtype=np.float32)
    nshell = np.ascontiguousarray(t["nshell"], dtype=np.int32)
    rho_nom = np.ascontiguousarray(t["density"], dtype=np.float32)
    mol = np.ascontiguousarray(mat_of_label, dtype=np.int32)
    _lib.check(_lib.load().drr_set_scatter_tables(projector._h, len(names), len(e), float(e[0]), float(e[1] - e[0]), _lib.ptr(mfp), _lib.ptr(rita),
                                                  _lib.ptr(comp), _lib.ptr(nshell), _lib.ptr(rho_nom), _lib.ptr(mol), _lib.ptr(rho_max)), projector._h)


def simulate(projector, proj, n_photons: int, seed: int = 0, photon_offset: int = 0, sdd: Optional[float] = None, out=None):
    """Photons [photon_offset, photon_offset + n_photons) for one view -> (tally uint64 [H, W], counters float64 [8]).

    ``out``: an int64 CUDA tensor [H, W] on the projector's GPU that receives the tally instead of a host array (the tally then
    stays in device memory, e.g. for the NCCL all-reduce over the ranks that share a view's photons)."""
    from . import geo

    sdd = float(sdd if sdd is not None else projector.source_to_detector_distance)
    if not sdd > 0:
        raise ValueError("scatter needs the source-to-detector distance (pass a device)")
    W, H = proj.intrinsic.sensor_size
    w2i, _, ijk = geo.pose_arrays(proj, projector.volumes)
    p_idx = np.ascontiguousarray(np.asarray(proj.index_from_world, dtype=np.float64)[:3, :] / sdd, dtype=np.float32)
    src = np.ascontiguousarray(np.asarray(proj.center_in_world, dtype=np.float64).reshape(-1)[:3], dtype=np.float32)
    counters = np.zeros(8, dtype=np.float64)
    V = len(projector.volumes)
    pr = np.ascontiguousarray(projector.priorities, dtype=np.int32)
    en = np.ascontiguousarray([1 if getattr(v, "enabled", True) else 0 for v in projector.volumes], dtype=np.int32)
    _lib.check(_lib.load().drr_set_priorities(projector._h, _lib.ptr(pr), _lib.ptr(en), V), projector._h)
    if out is not None:
        if not (hasattr(out, "is_cuda") and out.is_cuda and out.is_contiguous() and out.numel() == W * H and out.element_size() == 8):
            raise ValueError("out must be a contiguous 64-bit integer CUDA tensor of H x W elements")
        tally, mem = out, _lib.MEM_DEVICE
    else:
        tally, mem = np.zeros((H, W), dtype=np.uint64), _lib.MEM_HOST
    _lib.check(_lib.load().drr_scatter(projector._h, int(n_photons), int(photon_offset), int(seed) & 0xFFFFFFFFFFFFFFFF, W, H, _lib.ptr(w2i),
                                       _lib.ptr(p_idx), _lib.ptr(src), _lib.ptr(np.ascontiguousarray(ijk)), _lib.ptr(tally), _lib.ptr(counters),
                                       mem), projector._h)
    return tally, counters


def _dist():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except Exception:
        pass
    return None


def simulate_sharded(projector, proj, n_photons_total: int, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """One view's photons over all ranks of the default process group: rank r simulates ids [r n / w, (r + 1) n / w) of the same
    Philox stream, the fixed-point tallies are summed with ONE all-reduce -- on NCCL the tally goes kernel -> device tensor ->
    all-reduce over NVLink -> one copy to the host, never through host memory in between -- and every rank returns the full
    (tally uint64 [H, W], this rank's counters).  Without a process group this is ``simulate``."""
    from .parallel import shard_range

    dist = _dist()
    if dist is None:
        return simulate(projector, proj, n_photons_total, seed=seed)
    import torch

    a, b = shard_range(int(n_photons_total), dist.get_rank(), dist.get_world_size())
    if dist.get_backend() == "nccl":
        W, H = proj.intrinsic.sensor_size
        t = torch.zeros((H, W), dtype=torch.int64, device=torch.device("cuda", int(projector.cuda_device_id or 0)))
        _, counters = simulate(projector, proj, b - a, seed=seed, photon_offset=a, out=t)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy().view(np.uint64), counters
    tally, counters = simulate(projector, proj, b - a, seed=seed, photon_offset=a)
    return reduce_over_ranks(tally), counters


def reduce_over_ranks(tally: np.ndarray) -> np.ndarray:
    """Sum host-side integer tallies of all ranks (gloo on CPU; on NCCL prefer ``simulate_sharded``, which keeps the tally on the
    device); identity without torch.distributed."""
    dist = _dist()
    if dist is None:
        return tally
    import torch

    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(tally.view(np.int64).copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().view(np.uint64)


def scatter_image(tally: np.ndarray, n_photons_total: int, proj) -> np.ndarray:
    """Scatter signal in the units of the primary image: keV arriving in a pixel per photon emitted towards it.

    Photons are emitted uniformly over the detector area with weight |r|^-3 (solid angle), r = world_from_index (u, v, 1);
    the expected weighted number emitted towards pixel p is N / (W H) * |r_p|^-3.
    """
    H, W = tally.shape
    m = np.asarray(proj.world_from_index, dtype=np.float64)[:3, :]
    u, v = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    r = np.stack([u, v, np.ones_like(u)], axis=-1) @ m.T
    wgt = np.linalg.norm(r, axis=-1) ** -3
    emitted = n_photons_total / float(W * H) * wgt
    return (tally.astype(np.float64) / 65536.0 / 1000.0 / emitted).astype(np.float32)
