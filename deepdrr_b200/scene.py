"""Host-side scene bookkeeping shared by the Projector, the tests and the bench.

Follows the reference's constructor / ``initialize`` logic:

* material universe = sorted set of every volume's material names (+ mesh material names, + "air"
  when attenuating outside the volume)                         (projector.py:547-559)
* per-volume label remap ``remap[k] = sorted_names.index(name_k)`` in dict order, cast to uint8
                                                                (projector.py:1499-1509)
* default priorities ``[N-1, ..., 0]`` (later volumes win)      (projector.py:489-492)
* spectrum tables (keV energies, normalised pdf) and the ``[bin * M + m]`` mu/rho table
                                                                (projector.py:1659-1686)
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .material import absorb_coef_table
from .spectral_data import get_spectrum, spectrum_tables


def material_universe(volumes: Sequence, mesh_materials: Sequence[str] = (), attenuate_outside_volume: bool = False) -> List[str]:
    all_mats: List[str] = []
    for v in volumes:
        all_mats.extend(list(v.materials[0].keys()))
    all_mats.extend(mesh_materials)
    if attenuate_outside_volume:
        all_mats.append("air")
    out = list(set(all_mats))
    out.sort()
    return out


def remap_labels(volume, all_materials: Sequence[str]) -> np.ndarray:
    """uint8 ``[Ni, Nj, Nk]`` labels in the global material index (projector.py:1499-1509)."""
    label_list = [all_materials.index(k) for k in volume.materials[0] if k in all_materials]
    remap = np.array(label_list, dtype=np.uint16)
    return remap[volume.materials[1]].astype(np.uint8)


def default_priorities(n: int) -> List[int]:
    return [n - 1 - i for i in range(n)]


class SceneTables:
    """Everything the kernels need that does not depend on the view."""

    def __init__(self, volumes: Sequence, spectrum="90KV_AL40", mesh_materials: Sequence[str] = (),
                 attenuate_outside_volume: bool = False, priorities: Optional[Sequence[int]] = None):
        self.volumes = list(volumes)
        self.all_materials = material_universe(volumes, mesh_materials, attenuate_outside_volume)
        self.M = len(self.all_materials)
        self.labels = [remap_labels(v, self.all_materials) for v in volumes]
        self.priorities = list(priorities) if priorities is not None else default_priorities(len(volumes))
        self.spectrum_arr = get_spectrum(spectrum)
        self.energies, self.pdf = spectrum_tables(self.spectrum_arr)
        self.mu = absorb_coef_table(self.all_materials, self.energies)
        self.n_bins = len(self.energies)
