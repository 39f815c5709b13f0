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


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def random_single_volume_scene(rng, it, build=True):
    """Scene number ``it`` of the randomised parity runs (tools/fuzz_single.py): a thorax phantom of random shape, spacing and pose
    seen by a random camera -- source inside the volume (it % 4 == 0), grazing along a face (1) or anywhere around it (2, 3), coarse
    to fine detectors, short or unbounded ``max_ray_length``.  Consumes the same random numbers whether or not the volume is built,
    so that one iteration of a long run can be replayed."""
    from deepdrr_b200 import geo

    shape = tuple(int(x) for x in rng.integers(24, 72, size=3))
    spacing = tuple(rng.uniform(0.6, 8.0, size=3))
    vseed = int(rng.integers(1 << 30))
    vrot, vtr = _random_rotation(rng), rng.uniform(-40, 40, size=3)
    v = st = None
    if build:
        v = phantoms.thorax_volume(shape, spacing, seed=vseed)
        v.rotate(vrot)
        v.translate(vtr)
        st = SceneTables([v], "90KV_AL40")
    W, H = int(rng.integers(17, 120)), int(rng.integers(9, 100))
    pixel = float(rng.choice([0.2, 0.8, 2.0, 6.0]))
    sdd = float(rng.uniform(300, 1500))
    k = geo.CameraIntrinsicTransform.from_sizes((W, H), pixel, sdd)
    extent = np.array(shape) * np.array(spacing)
    mode = it % 4
    if mode == 0:    # source inside the volume
        source = rng.uniform(-0.3, 0.3, size=3) * extent
    elif mode == 1:  # grazing: looking along a face
        source = np.array([0.0, -extent[1], 0.5 * extent[2]]) + rng.normal(size=3)
    else:
        source = rng.normal(size=3)
        source = source / np.linalg.norm(source) * float(rng.uniform(0.6, 3.0)) * extent.max()
    direction = -source + rng.normal(size=3) * 0.2 * extent.max() if mode != 0 else rng.normal(size=3)
    up = rng.normal(size=3)
    pose = phantoms.look_at_projection(source, direction, up, k)
    mrl = float(rng.choice([4 * sdd, 0.7 * np.linalg.norm(source) + 10.0, 1e5]))
    return {"volume": v, "tables": st, "W": W, "H": H, "pixel": pixel, "sdd": sdd, "k": k, "pose": pose, "mrl": mrl, "mode": mode, "shape": shape}
