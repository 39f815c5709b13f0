#!/usr/bin/env python
"""Pack the MC-GPU photon-interaction tables the scatter kernel needs into deepdrr_b200/data/mcgpu_tables.npz.

Runs only in the build container (needs /root/reference).  The numbers are the MC-GPU material data files
(Badal & Badano 2009) that the reference ships as Python literals under deepdrr/projector/mcgpu_*:
mean free paths (mcgpu_mfp_data.py, cm), RITA Rayleigh form-factor sampling tables (mcgpu_rita_samplers.py),
Compton shell data (mcgpu_compton_data.py) and nominal densities (mcgpu_density.py).  Physical data, not code.
The 5 eV energy grid is thinned to 100 eV (the mean free paths are smooth above 5 keV; the kernel interpolates
log-linearly), which takes 10.5 MB down to 0.5 MB.
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("DEEPDRR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pkg = types.ModuleType("deepdrr"); pkg.__path__ = [os.path.join(REF, "deepdrr")]; sys.modules["deepdrr"] = pkg
prj = types.ModuleType("deepdrr.projector"); prj.__path__ = [os.path.join(REF, "deepdrr", "projector")]; sys.modules["deepdrr.projector"] = prj
from deepdrr.projector.mcgpu_compton_data import COMPTON_DATA, MATERIAL_NSHELLS  # noqa: E402
from deepdrr.projector.mcgpu_density import density_data  # noqa: E402
from deepdrr.projector.mcgpu_mfp_data import MFP_DATA  # noqa: E402

import importlib  # noqa: E402

names = sorted(MFP_DATA.keys())
thin = 20
mfp = np.stack([np.asarray(MFP_DATA[n], dtype=np.float64)[::thin] for n in names])       # [n_mat, 1151, 6], mm
assert np.allclose(mfp[:, :, 0], mfp[0, :, 0])
rita = []
mod = importlib.import_module("deepdrr.projector.mcgpu_rita_samplers")
for n in names:
    key = [k for k in dir(mod) if k.endswith("_RITA_PARAMS") and k.lower().startswith(n.lower().replace(" ", "_").split("_")[0])]
    cand = [k for k in key if n.lower().replace(" ", "_") in k.lower() or n.lower().split(" ")[0] in k.lower()]
    arr = np.asarray(getattr(mod, cand[0]), dtype=np.float64)
    rita.append(arr[:, :4])
    print(n, "<-", cand[0], arr.shape)
rita = np.stack(rita)                                                                      # [n_mat, 128, 4] = x^2, P, A, B
compton = np.zeros((len(names), 30, 3), dtype=np.float64)
nshell = np.zeros(len(names), dtype=np.int32)
for i, n in enumerate(names):
    c = np.asarray(COMPTON_DATA[n], dtype=np.float64)
    nshell[i] = int(MATERIAL_NSHELLS[n])
    compton[i, : nshell[i]] = c[: nshell[i], :3]                                           # FCO (electrons), UICO (eV), FJ0
rho = np.array([density_data[n] for n in names], dtype=np.float64)
np.savez_compressed(os.path.join(ROOT, "deepdrr_b200", "data", "mcgpu_tables.npz"), names=np.array(names), energy_eV=mfp[0, :, 0].astype(np.float32),
                    mfp_mm=mfp[:, :, 1:6].astype(np.float32), rita=rita.astype(np.float32), compton=compton.astype(np.float32), nshell=nshell,
                    density=rho.astype(np.float32))
print("materials:", names, "energies:", mfp.shape[1], os.path.getsize(os.path.join(ROOT, "deepdrr_b200", "data", "mcgpu_tables.npz")) / 1e6, "MB")
