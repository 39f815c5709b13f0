"""TEST INFRASTRUCTURE -- ctypes wrapper of oracle/scatter_oracle.c (CPU restatement of the scatter transport).

Builds the same inputs ``deepdrr_b200.scatter.setup`` / ``simulate`` hand to ``drr_set_scatter_tables`` / ``drr_scatter`` (tables
from data/mcgpu_tables.npz, majorant, spectrum CDF, pose matrices) and runs the plain-C transport on them.  Never imported by the
product.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libscatter_oracle.so")
_lib = None


class _Scene(ctypes.Structure):
    _fields_ = [("n_mat", ctypes.c_int), ("n_e", ctypes.c_int), ("e0", ctypes.c_float), ("de", ctypes.c_float),
                ("mfp", ctypes.c_void_p), ("rita", ctypes.c_void_p), ("compton", ctypes.c_void_p), ("nshell", ctypes.c_void_p),
                ("inv_rho_nom", ctypes.c_void_p), ("majorant", ctypes.c_void_p), ("mat_of_label", ctypes.c_void_p), ("s0", ctypes.c_void_p),
                ("V", ctypes.c_int), ("priority", ctypes.c_int * 8), ("enabled", ctypes.c_int * 8),
                ("dens", ctypes.c_void_p * 8), ("lab", ctypes.c_void_p * 8), ("shape", (ctypes.c_int * 3) * 8),
                ("ijk", (ctypes.c_float * 12) * 8), ("p_idx", ctypes.c_float * 12), ("w2i", ctypes.c_float * 9), ("src", ctypes.c_float * 3),
                ("W", ctypes.c_int), ("H", ctypes.c_int), ("n_bins", ctypes.c_int), ("spec_e_keV", ctypes.c_void_p), ("spec_cdf", ctypes.c_void_p)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "scatter_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.drr_scatter_oracle_scene_size.restype = ctypes.c_size_t
        assert lib.drr_scatter_oracle_scene_size() == ctypes.sizeof(_Scene), "sc_scene layout mismatch"
        lib.drr_scatter_oracle.argtypes = [ctypes.POINTER(_Scene), ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        _lib = lib
    return _lib


def compton_samples(material: str, energy_eV: float, n: int, seed: int = 0):
    """(cos(theta) [n], E' [n]) of n Compton events in a table material (unit test of the impulse-approximation sampler)."""
    from deepdrr_b200.scatter import load_tables

    t = load_tables()
    names = [str(x) for x in t["names"]]
    comp = np.ascontiguousarray(t["compton"], dtype=np.float32)
    nshell = np.ascontiguousarray(t["nshell"], dtype=np.int32)
    S = _Scene()
    S.n_mat = len(names)
    S.compton, S.nshell = comp.ctypes.data, nshell.ctypes.data
    cost, e_out = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.float32)
    lib = _load()
    lib.drr_compton_oracle.argtypes = [ctypes.POINTER(_Scene), ctypes.c_int, ctypes.c_float, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    assert lib.drr_compton_oracle(ctypes.byref(S), names.index(material), float(energy_eV), int(seed), int(n), cost.ctypes.data, e_out.ctypes.data) == 0
    return cost, e_out


def simulate(volumes, all_materials, spectrum_energies_keV, spectrum_pdf, proj, sdd: float, n_photons: int, seed: int = 0, photon_offset: int = 0,
             priorities=None):
    """Same contract as ``deepdrr_b200.scatter.simulate``: (tally uint64 [H, W], counters float64 [8]).  ``volumes``: one Volume or a list."""
    from deepdrr_b200 import geo
    from deepdrr_b200.scatter import MCGPU_NAME, load_tables
    from deepdrr_b200.scene import default_priorities, remap_labels

    volumes = list(volumes) if isinstance(volumes, (list, tuple)) else [volumes]
    V = len(volumes)
    priorities = list(priorities) if priorities is not None else default_priorities(V)
    t = load_tables()
    names = [str(n) for n in t["names"]]
    mol = np.ascontiguousarray([names.index(MCGPU_NAME[m]) for m in all_materials], dtype=np.int32)
    labels = [np.ascontiguousarray(remap_labels(v, all_materials)) for v in volumes]
    dens = [np.ascontiguousarray(v.data, dtype=np.float32) for v in volumes]
    rho_max = np.zeros(len(all_materials), dtype=np.float32)
    for d, lab in zip(dens, labels):
        for l in range(len(all_materials)):
            if np.any(lab == l):
                rho_max[l] = max(rho_max[l], np.float32(d[lab == l].max()))
    e = t["energy_eV"].astype(np.float64)
    mfp = np.ascontiguousarray(t["mfp_mm"], dtype=np.float32)
    rita = np.ascontiguousarray(t["rita"], dtype=np.float32)
    comp = np.ascontiguousarray(t["compton"], dtype=np.float32)
    nshell = np.ascontiguousarray(t["nshell"], dtype=np.int32)
    inv_rho = (np.float32(1.0) / np.ascontiguousarray(t["density"], dtype=np.float32)).astype(np.float32)
    # majorant: max over the scene's materials of rho_max / rho_nom / mfp_total, float32 as in drr_set_scatter_tables
    maj = np.zeros(len(e), dtype=np.float32)
    for l, m in enumerate(mol):
        mu = (rho_max[l] * inv_rho[m]).astype(np.float32) / mfp[m, :, 3]
        maj = np.maximum(maj, mu.astype(np.float32))
    maj[~(maj > 0)] = np.float32(1e-6)
    # S(E, theta = pi) of every table material on the energy grid, x 1.001 (float64, then float32): the rejection's normalisation
    s0 = np.zeros((len(names), len(e)), dtype=np.float64)
    REV, D2, D1 = 510998.918, 1.4142135623731, 0.70710678118655
    for m in range(len(names)):
        for i in range(int(nshell[m])):
            f, U, J = (float(x) for x in comp[m, i])
            on = U < e
            aux = e * (e - U) * 2.0
            with np.errstate(invalid="ignore"):
                pz = J * (aux - REV * U) / (REV * np.sqrt(aux + aux + U * U))
            q = np.where(pz > 0, D1 + D2 * pz, D1 - D2 * pz)
            h = 0.5 * np.exp(0.5 - q * q)
            s0[m] += np.where(on, f * np.where(pz > 0, 1.0 - h, h), 0.0)
    s0 = np.ascontiguousarray((s0 * 1.001).astype(np.float32))
    pdf = np.asarray(spectrum_pdf, dtype=np.float32)
    pos = np.where(pdf > 0, pdf.astype(np.float64), 0.0)
    cdf = (np.cumsum(pos) / pos.sum()).astype(np.float32)
    cdf[-1] = 1.0
    ekev = np.ascontiguousarray(spectrum_energies_keV, dtype=np.float32)
    W, H = proj.intrinsic.sensor_size
    w2i, _, ijk = geo.pose_arrays(proj, volumes)
    p_idx = np.ascontiguousarray(np.asarray(proj.index_from_world, dtype=np.float64)[:3, :] / float(sdd), dtype=np.float32).reshape(12)
    src = np.ascontiguousarray(np.asarray(proj.center_in_world, dtype=np.float64).reshape(-1)[:3], dtype=np.float32)
    S = _Scene()
    S.n_mat, S.n_e, S.e0, S.de = len(names), len(e), float(e[0]), float(e[1] - e[0])
    keep = [mfp, rita, comp, nshell, inv_rho, maj, mol, dens, labels, ekev, cdf, s0]
    S.mfp, S.rita, S.compton, S.nshell = mfp.ctypes.data, rita.ctypes.data, comp.ctypes.data, nshell.ctypes.data
    S.inv_rho_nom, S.majorant, S.mat_of_label = inv_rho.ctypes.data, maj.ctypes.data, mol.ctypes.data
    S.s0 = s0.ctypes.data
    S.V = V
    for v in range(V):
        S.priority[v], S.enabled[v] = int(priorities[v]), int(bool(getattr(volumes[v], "enabled", True)))
        S.dens[v], S.lab[v] = dens[v].ctypes.data, labels[v].ctypes.data
        S.shape[v][:] = [int(x) for x in dens[v].shape]
        S.ijk[v][:] = [float(x) for x in ijk[v]]
    S.p_idx[:] = [float(x) for x in p_idx]
    S.w2i[:] = [float(x) for x in w2i]
    S.src[:] = [float(x) for x in src]
    S.W, S.H, S.n_bins = int(W), int(H), len(ekev)
    S.spec_e_keV, S.spec_cdf = ekev.ctypes.data, cdf.ctypes.data
    tally = np.zeros((H, W), dtype=np.uint64)
    counters = np.zeros(8, dtype=np.float64)
    rc = _load().drr_scatter_oracle(ctypes.byref(S), int(n_photons), int(photon_offset), int(seed) & 0xFFFFFFFFFFFFFFFF, tally.ctypes.data,
                                    counters.ctypes.data)
    assert rc == 0
    del keep
    return tally, counters
