"""The mesh path against the reference's OWN golden image: tests/reference/test_mesh_mesh_1.gif is what the reference's renderer
(pyrender + GLSL dual depth peeling + kernelTide + projectKernel + neglog) produced for its mesh-only test
(tests/test_core.py:353-470): an additive titanium body, four density-0 subtractive cubes and a 10 000 x scaled subtractive
threaded rod on layer 1, a SimpleDevice turning about the scene.  No CT is involved, so the test can be re-staged offline: the
STLs are in tests/golden/mesh_fixtures.npz, six of the twenty truth frames in tests/golden/ref_test_mesh_mesh_1.npz
(tools/gen_mesh_fixtures.py).  The reference compares 8-bit images (`verify_image`); so does this test, with a small allowance
for silhouette pixels, where a rasteriser and a ray tracer decide coverage differently.
"""
import os

import numpy as np
import pytest

import cases
from deepdrr_b200 import Projector, geo
from deepdrr_b200.device import SimpleDevice
from deepdrr_b200.vol import Mesh

FIX = os.path.join(cases.GOLDEN, "mesh_fixtures.npz")
TRUTH = os.path.join(cases.GOLDEN, "ref_test_mesh_mesh_1.npz")


def _mesh(name, scale=1.0, **kw):
    tris = np.load(FIX)[name].astype(np.float32) * np.float32(scale)
    return Mesh(tris.reshape(-1, 3), np.arange(tris.shape[0] * 3).reshape(-1, 3), **kw)


def _rot(axis, angle):
    c, s = np.cos(angle), np.sin(angle)
    m = {"x": [[1, 0, 0], [0, c, -s], [0, s, c]], "y": [[c, 0, s], [0, 1, 0], [-s, 0, c]]}[axis]
    return geo.FrameTransform.from_rt(np.array(m, dtype=np.float64))


def render_frames(frame_ids, N=20):
    """tests/test_core.py:353-470, statement by statement (same object order, materials, layers, motion and camera)."""
    cubes = [_mesh(f"mm1_cube{i}", material="bone", density=0.0, subtractive=True, layer=1,
                   world_from_anatomical=geo.FrameTransform.from_translation([10, 30, 5])) for i in (1, 2, 3, 4)]          # :369-372
    body = _mesh("mm1_body", material="titanium", density=0.1, subtractive=False,
                 world_from_anatomical=geo.FrameTransform.from_translation([0, 20, 0]))                                      # :374-375
    rod = _mesh("threads", 10000.0, material="titanium", density=0.0, subtractive=True, layer=1)                             # :380-392
    carm = SimpleDevice(sensor_width=400, sensor_height=400, pixel_size=8.0, source_to_detector_distance=4000)               # :399
    rand_coeffs = np.random.RandomState(2).rand(4, 3) * 100                                                                  # :426-427
    out = []
    with Projector([body, rod] + cubes, device=carm, step=0.01, mode="linear", max_block_index=65535, spectrum="90KV_AL40", photon_count=100000,
                   scatter_num=0, threads=8, max_mesh_hits=128) as projector:                                                # :410-421
        for i in frame_ids:
            for m_idx, m in enumerate(cubes):                                                                                # :446-466
                a = geo.FrameTransform.from_translation([300 * np.sin(i / N * np.pi * 2 * 3 + rand_coeffs[m_idx, 0]), 0,
                                                         300 * np.sin(i / N * np.pi * 2 * 3 + rand_coeffs[m_idx, 2])])
                if m_idx == 2:
                    rod.world_from_anatomical = a
                if m_idx in (2, 1):
                    a = geo.FrameTransform.from_translation([0, 0, 0])
                m.world_from_anatomical = a
            new = _rot("x", -np.pi / 2) @ _rot("y", -i / N * np.pi * 2) @ geo.FrameTransform.from_translation([0, 0, -2000])  # :468-473
            carm._device_from_camera3d = new                                                                                  # :475
            out.append(np.array(projector.project()))                                                                         # :477
    return out


@pytest.mark.gpu
def test_mesh_only_scene_matches_the_references_truth_frames():
    g = np.load(TRUTH)
    ids = [int(i) for i in g["frame_ids"]]
    imgs = render_frames(ids)
    for n, i in enumerate(ids):
        want = g["frames"][n].astype(np.int32)
        got = (imgs[n] * 255).astype(np.uint8).astype(np.int32)                                                               # :479
        assert got.shape == want.shape == (400, 400)
        d = np.abs(got - want)
        # measured on a B200: every pixel of all six frames within 1/255 (mean |d| 0.001 - 0.003); the allowance is for silhouettes
        assert np.mean(d <= 1) >= 0.9995, f"frame {i}: only {np.mean(d <= 1):.4f} of the pixels within 1/255"
        assert d.max() <= 3, f"frame {i}: a pixel is off by {d.max()}/255"
