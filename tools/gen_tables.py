#!/usr/bin/env python
"""Generate the packaged physics tables and the table golden vectors.

Runs ONLY in the build container (needs /root/reference).  It imports the reference's own
``deepdrr.material`` and ``deepdrr.projector.spectral_data`` through a stub parent package
(SURVEY.md App. D probe 2) and writes

* ``deepdrr_b200/data/nist_mu_rho.npz``  -- raw NIST (energy MeV, mu/rho, mu_en/rho) rows of every
  file under ``deepdrr/material/material_decompositions`` plus the name map
  (``deepdrr/material/mappings.py``).  These are NIST XCOM physical constants (data, not code).
* ``deepdrr_b200/data/spectra.npz``      -- the three spectra of ``spectral_data.py:463``.
* ``tests/golden/absorb_tables.npz``     -- absorb_coef_table / energies / pdf exactly as
  ``projector.py:1659-1686`` builds them with the reference's ``Material`` class; the CPU tests
  pin ``deepdrr_b200.material`` against these.
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("DEEPDRR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_import():
    pkg = types.ModuleType("deepdrr")
    pkg.__path__ = [os.path.join(REF, "deepdrr")]
    sys.modules["deepdrr"] = pkg
    prj = types.ModuleType("deepdrr.projector")
    prj.__path__ = [os.path.join(REF, "deepdrr", "projector")]
    sys.modules["deepdrr.projector"] = prj
    from deepdrr.material import Material  # noqa
    from deepdrr.projector import spectral_data  # noqa

    return Material, spectral_data


def main():
    Material, spectral_data = _stub_import()
    mdir = os.path.join(REF, "deepdrr", "material", "material_decompositions")
    out = {}
    names = sorted(os.listdir(mdir))
    for n in names:
        m = Material.from_string(n)
        out["tab::" + n] = np.stack([m.energy, m.mu_over_rho, m.mu_en_over_rho], axis=1).astype(np.float64)
    cmap = dict(Material._custom_map)
    out["map_keys"] = np.array(list(cmap.keys()))
    out["map_vals"] = np.array([cmap[k] for k in cmap])
    np.savez_compressed(os.path.join(ROOT, "deepdrr_b200", "data", "nist_mu_rho.npz"), **out)

    sp = {k: np.asarray(v, dtype=np.float64) for k, v in spectral_data.spectrums.items()}
    np.savez_compressed(os.path.join(ROOT, "deepdrr_b200", "data", "spectra.npz"), **sp)

    # golden absorb tables (projector.py:1659-1686)
    gold = {}
    mats = ["air", "bone", "iron", "lung", "soft tissue", "titanium", "blood", "muscle", "Au", "H"]
    gold["materials"] = np.array(mats)
    for sname, arr in sp.items():
        energies = np.ascontiguousarray(arr[:, 0].copy() / 1000, dtype=np.float32)
        pdf = np.ascontiguousarray((arr[:, 1] / np.sum(arr[:, 1])).copy(), dtype=np.float32)
        tab = np.zeros(len(energies) * len(mats)).astype(np.float32)
        for b in range(len(energies)):
            for m, mn in enumerate(mats):
                tab[b * len(mats) + m] = Material.from_string(mn).get_coefficients(energies[b]).mu_over_rho
        gold[sname + "::energies"] = energies
        gold[sname + "::pdf"] = pdf
        gold[sname + "::table"] = tab
    # one compound-string material as well (material.py:141-147)
    cs = "H0.111900O0.888100"
    m = Material.from_string(cs, compound_string=True)
    gold["compound::name"] = np.array(cs)
    gold["compound::mu60"] = np.array(m.get_coefficients(60.0).mu_over_rho)
    gold["compound::mu33"] = np.array(m.get_coefficients(33.3).mu_over_rho)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "absorb_tables.npz"), **gold)
    print("wrote tables:", len(names), "materials;", list(sp))


if __name__ == "__main__":
    main()
