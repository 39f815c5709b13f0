"""Shared deterministic test scenes (same recipes as tools/make_goldens.py used on the GPU box)."""
import os

import numpy as np

from deepdrr_b200 import phantoms
from deepdrr_b200.scene import SceneTables

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def scene(name):
    """Returns (volumes, spectrum, priorities) of a golden case."""
    if name == "c1":
        return [phantoms.c1_volume()], "90KV_AL40", None
    vs = phantoms.thorax_volume((128, 128, 100), (3.2, 3.2, 4.0))
    if name == "thorax_small":
        return [vs], "120KV_AL43", None
    if name == "multivol3":
        w = phantoms.kwire_volume(length_mm=60.0, spacing=0.25, half_width=6)
        phantoms.place_kwire(w, (-10.0, -20.0, 0.0), (0.3, 1.0, 0.1))
        w2 = phantoms.kwire_volume(length_mm=60.0, spacing=0.25, half_width=6)
        phantoms.place_kwire(w2, (10.0, -20.0, 5.0), (-0.3, 1.0, 0.0))
        return [vs, w, w2], "90KV_AL40", None
    if name == "multivol2_sameprio":
        vs2 = phantoms.thorax_volume((96, 96, 80), (3.2, 3.2, 4.0), seed=3)
        vs2.translate((15.0, -10.0, 20.0))
        return [vs, vs2], "60KV_AL35", [0, 0]
    if name == "c2":
        return [phantoms.thorax_volume()], "120KV_AL43", None
    raise KeyError(name)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def n_views(g):
    i = 0
    while f"w2i_{i}" in g:
        i += 1
    return i


def tables(volumes, spectrum, priorities):
    return SceneTables(volumes, spectrum, priorities=priorities)


def rel_err(a, b, floor=1e-30):
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(np.abs(b.astype(np.float64)), floor)


class MatrixProjection:
    """A 'camera projection' given directly by stored kernel matrices (identical inputs for both sides)."""

    def __init__(self, w2i, W, H):
        self.world_from_index = np.concatenate([np.asarray(w2i, dtype=np.float64).reshape(3, 3), np.zeros((1, 3))], axis=0)
        self.W, self.H = W, H

        class _K:
            sensor_size = (W, H)
            sensor_width, sensor_height = W, H
            fx = fy = 1.0

        self.intrinsic = _K()
