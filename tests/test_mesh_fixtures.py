"""The reference's STL fixtures (BASELINE config 4's screw data/6.5mmD_32mmThread_L130mm.STL; tests/resources/10cmcube.stl,
threads.stl, suzanne.stl) as the mesh tests use them: packed into tests/golden/mesh_fixtures.npz by tools/gen_mesh_fixtures.py
because /root/reference does not exist on the GPU box.  Here: the fixtures equal what ``Mesh.from_stl`` reads from the files
(wherever the files are present), the meshes are closed, and the CUDA ray-triangle path gives analytic chord lengths for the
cube and the float64 restatement's values for the real screw."""
import os

import numpy as np
import pytest

import cases
from deepdrr_b200 import Projector, geo, phantoms
from deepdrr_b200.vol import Mesh
from oracle import mesh_oracle

FIX = os.path.join(cases.GOLDEN, "mesh_fixtures.npz")
REF = "/root/reference"
FILES = {"screw": "data/6.5mmD_32mmThread_L130mm.STL", "cube": "tests/resources/10cmcube.stl", "threads": "tests/resources/threads.stl",
         "suzanne": "tests/resources/suzanne.stl"}


def _mesh(name, scale=1.0, **kw):
    tris = np.load(FIX)[name].astype(np.float32) * np.float32(scale)
    return Mesh(tris.reshape(-1, 3), np.arange(tris.shape[0] * 3).reshape(-1, 3), **kw)


def test_fixtures_equal_the_reference_files_and_are_closed():
    g = np.load(FIX)
    assert g["screw"].shape == (7806, 3, 3) and g["cube"].shape == (12, 3, 3)
    for name, rel in FILES.items():
        path = os.path.join(REF, rel)
        if os.path.exists(path):  # this container; absent on the GPU box
            assert np.array_equal(Mesh.from_stl(path, material="titanium").triangles, g[name]), name
    for name in ("screw", "cube"):
        # watertight: every undirected edge is shared by exactly two triangles (README.md:257 requires closed meshes)
        t = np.round(g[name].astype(np.float64), 4)
        _, inv = np.unique(t.reshape(-1, 3), axis=0, return_inverse=True)
        f = inv.reshape(-1, 3)
        e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
        _, counts = np.unique(e, axis=0, return_counts=True)
        assert np.all(counts == 2), name


@pytest.mark.gpu
def test_cube_chords_are_analytic():
    """reference tests/test_core.py:219-227 scales 10cmcube.stl by 100 - 200; here x500: a 100 mm cube, rotated like there."""
    cube = _mesh("cube", 500.0, material="iron", density=2.0)
    a = np.deg2rad(60.0)
    rot = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    cube.world_from_anatomical = geo.FrameTransform.from_rt(rot, (5.0, -3.0, 8.0))
    k = geo.CameraIntrinsicTransform.from_sizes((160, 128), 1.5, 1000.0)
    pose = phantoms.look_at_projection((30.0, -600.0, 40.0), (-0.05, 1.0, -0.07), (0, 0, 1), k)
    with Projector([cube], camera_intrinsics=k, source_to_detector_distance=1000.0, neglog=False) as p:
        area = p.project_line_integrals(pose)[0, 0]
    # analytic: slab intersection of every pixel ray with the box, in the cube's own frame
    w2i, _, _ = geo.pose_arrays(pose, [])
    dirs = mesh_oracle.pixel_dirs(w2i, 160, 128)
    inv = np.linalg.inv(cube.world_from_ijk.data)
    o = inv[:3, :3] @ pose.center_in_world + inv[:3, 3]
    d = dirs @ inv[:3, :3].T
    with np.errstate(divide="ignore", invalid="ignore"):
        t0, t1 = (-50.0 - o) / d, (50.0 - o) / d
    near, far = np.minimum(t0, t1).max(axis=1), np.maximum(t0, t1).min(axis=1)
    chord = np.where(far > near, far - near, 0.0).reshape(128, 160)
    want = chord * 2.0 / 10.0
    interior = chord > 0.5                       # not grazing an edge of the silhouette
    assert interior.sum() > 3000
    assert np.all(area[chord == 0] == 0) or (area[chord == 0] > 0).mean() < 0.002
    # hit distances are float32 (as in the reference's RG32F targets): 2 hits x a few ulps of ~600 mm
    assert np.abs(area - want)[interior].max() <= 1e-5 * want[interior].max() + 8 * 6.1e-5 * 2.0 / 10.0


@pytest.mark.gpu
def test_config4_real_screw_over_ct_matches_restatement():
    """BASELINE config 4: CT + the titanium screw of data/6.5mmD_32mmThread_L130mm.STL, 384^2 sensor.  The titanium line integral is
    checked against a float64 ray-triangle restatement of the GL semantics; every other material must equal
    the projection of the CT alone, since an additive mesh changes nothing in the march (K.cu:569-579)."""
    screw = _mesh("screw", 1.0, material="titanium")
    # the STL's long axis is its local y (0 .. 130 mm): tilt it in the detector plane of the views (which look along ~ +z) and put
    # the threaded end in the field of view (57 mm wide at the isocentre)
    ax = np.array([0.8, 0.55, 0.25]) / np.linalg.norm([0.8, 0.55, 0.25])
    e1 = np.cross(ax, [0.0, 0.0, 1.0]); e1 /= np.linalg.norm(e1)
    rot = np.stack([e1, ax, np.cross(e1, ax)], axis=1)                  # local x, y, z -> world
    tip_local = np.array([4.1, 0.0, 4.1])                               # on the screw's axis, at its y = 0 end
    screw.world_from_anatomical = geo.FrameTransform.from_rt(rot, np.array([-14.0, -9.0, 3.0]) - rot @ tip_local)
    ct = phantoms.thorax_volume((128, 128, 100), (3.2, 3.2, 4.0))
    poses, sdd = phantoms.cone_poses(2, seed=4)
    k = poses[0].intrinsic
    with Projector([ct, screw], spectrum="120KV_AL43", neglog=False, camera_intrinsics=k, source_to_detector_distance=sdd) as p:
        mats = p.all_materials
        area = p.project_line_integrals(*poses)
    with Projector([ct], spectrum="120KV_AL43", neglog=False, camera_intrinsics=k, source_to_detector_distance=sdd) as p:
        mats0 = p.all_materials
        area0 = p.project_line_integrals(*poses)
    ti = mats.index("titanium")
    for m0, name in enumerate(mats0):
        # an additive mesh adds its own term after the march (K.cu:569-579) and leaves the CT's samples alone; the scene with a mesh
        # runs the multi-object kernel, whose texture / FMA-pipe sampler mix differs from the single-volume kernel's, so the two
        # agree to the samplers' common distance from the reference (5e-7 each), not to the bit
        a, b = area[:, mats.index(name)], area0[:, m0]
        assert np.array_equal(a == 0, b == 0), name
        assert cases.rel_err(a, b)[b > 0].max() <= 2e-6, name
    tris_world = (screw.triangles.astype(np.float64).reshape(-1, 3) @ screw.world_from_ijk.data[:3, :3].T + screw.world_from_ijk.data[:3, 3]).reshape(-1, 3, 3)
    rho = screw.density
    for n, pose in enumerate(poses):
        w2i, _, _ = geo.pose_arrays(pose, [])
        dirs = mesh_oracle.pixel_dirs(w2i, 384, 384)
        src = pose.center_in_world
        # only the rays whose pixel lies in the screen box of the projected vertices can hit (checked: the others are zero)
        P = pose.index_from_world
        vh = tris_world.reshape(-1, 3) @ P[:, :3].T + P[:, 3]
        uv = vh[:, :2] / vh[:, 2:3]
        u0, v0 = np.floor(uv.min(axis=0)).astype(int) - 2
        u1, v1 = np.ceil(uv.max(axis=0)).astype(int) + 2
        uu, vv = np.meshgrid(np.arange(384), np.arange(384))
        inbox = ((uu >= u0) & (uu <= u1) & (vv >= v0) & (vv <= v1)).reshape(-1)
        got = area[n, ti].reshape(-1)
        assert np.all(got[~inbox] == 0)
        idx = np.nonzero(inbox)[0]
        R, G, cnt = np.zeros(len(idx)), np.zeros(len(idx)), np.zeros(len(idx), dtype=int)
        for a in range(0, len(idx), 256):                          # chunks keep the [rays, triangles] temporaries small
            t, ent = mesh_oracle.trace(tris_world, src, dirs[idx[a:a + 256]])
            fin = np.isfinite(t)
            s = np.where(ent, -1.0, 1.0)
            R[a:a + 256] = np.where(fin, np.where(fin, t, 0.0) * s * rho, 0.0).sum(axis=1)
            G[a:a + 256] = np.where(fin, s, 0.0).sum(axis=1)
            cnt[a:a + 256] = fin.sum(axis=1)
        want = np.where(np.abs(G) < 1e-5, np.maximum(R, 0.0), 0.0) / 10.0           # K.cu:569-584
        g = got[idx]
        # silhouette pixels: fp32 and fp64 edge tests disagree on the hit count there; everywhere else the chords must agree
        bad = np.abs(g - want) > 1e-3 * np.maximum(want, 0.05)
        assert bad.mean() < 0.03, bad.mean()
        ok = ~bad & (want > 0)
        assert ok.sum() > 2000, "the screw must be in view"
        atol = 8 * 6.1e-5 * rho / 10.0                                                # float32 hit distances at ~500-1000 mm
        assert ((np.abs(g - want) - atol)[ok] / want[ok]).max() <= 1e-5
        assert int(cnt.max()) >= 4, "a threaded screw has rays with more than one entry / exit pair"
