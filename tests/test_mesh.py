"""Mesh path (BASELINE config 4): CUDA ray-triangle hit intervals vs the NumPy restatement of the
reference's GL semantics (oracle/mesh_oracle.py) and vs the reference's own kernelTide.

The GL rasteriser itself cannot run offline, so parity is defined on what it samples: hit distances at
pixel centres.  Pixels on a silhouette (the two sides disagree on the hit count because of fp32 vs fp64
edge tests) are excluded; everywhere else area densities must agree to 1e-5 relative plus an absolute term
of 8 float32 ulps of the hit distance (6.1e-5 mm at 512-1024 mm from the source) times the densest mesh of
that material: hit distances are float32 here and in the reference's RG32F render targets alike, which
bounds how well a short chord through a dense mesh can be known by either.
"""
import numpy as np
import pytest

import cases
from deepdrr_b200 import Projector, geo, phantoms
from deepdrr_b200.vol import Mesh
from oracle import cpu_oracle, mesh_oracle


def _random_hit_lists(n_rays, n, rng):
    ts = np.full((n_rays, n), np.inf, dtype=np.float32)
    fs = np.zeros((n_rays, n), dtype=np.int8)
    for r in range(n_rays):
        k = rng.integers(0, n + 1)
        t = np.sort(rng.uniform(0.0, 2100.0, size=k)).astype(np.float32)
        if k > 2 and rng.random() < 0.3:
            t[1] = t[0]                                   # exact duplicates
        f = rng.choice(np.array([-1, 1], dtype=np.int8), size=k)
        if rng.random() < 0.5:                            # mostly well-formed entry/exit alternation
            f = np.where(np.arange(k) % 2 == 0, 1, -1).astype(np.int8)
        p = rng.permutation(k)
        ts[r, :k], fs[r, :k] = t[p], f[p]
    return ts, fs


def test_tide_restatement_properties():
    rng = np.random.default_rng(0)
    ts, fs = _random_hit_lists(200, 32, rng)
    for r in range(200):
        ct, cf = mesh_oracle.tide_clean(ts[r], fs[r], 2000.0)
        k = int((cf != 0).sum())
        assert np.all(cf[:k] != 0) and np.all(cf[k:] == 0) and np.all(np.isinf(ct[k:]))     # compacted
        assert np.all(np.diff(ct[:k]) >= 0)                                                  # sorted
        alt = np.cumsum(cf[:k].astype(int))
        assert np.all(alt >= 0) and np.all(alt <= 1)                                         # depth stays in {0, 1}


@pytest.mark.gpu
def test_cuda_tide_matches_restatement_and_reference_kernelTide():
    import ctypes
    from deepdrr_b200 import _lib
    from oracle import ref_gpu

    rng = np.random.default_rng(1)
    n_rays, n = 3000, 32
    ts, fs = _random_hit_lists(n_rays, n, rng)
    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.drr_create(0, ctypes.byref(h)))
    ct, cf = ts.copy(), fs.copy()
    _lib.check(lib.drr_mesh_clean_hits(h, _lib.ptr(ct), _lib.ptr(cf), n_rays, n, 2000.0, _lib.MEM_HOST), h)
    lib.drr_destroy(h)
    for r in range(0, n_rays, 7):
        et, ef = mesh_oracle.tide_clean(ts[r], fs[r], 2000.0)
        assert np.array_equal(ct[r], et) and np.array_equal(cf[r], ef)
    if not ref_gpu.available():
        pytest.skip("oracle/_ref not shipped")
    # the reference's own kernel: pack the same hits into its peel layout (-exit, +exit, -entry, +entry per pass)
    peel = np.zeros((n_rays, 32), dtype=np.float32)
    mine_t, mine_f = np.full((n_rays, 32), np.inf, np.float32), np.zeros((n_rays, 32), np.int8)
    for r in range(n_rays):
        ex = [t for t, f in zip(ts[r], fs[r]) if f == -1][:16]
        en = [t for t, f in zip(ts[r], fs[r]) if f == 1][:16]
        for i, t in enumerate(ex):
            peel[r, 4 * (i // 2) + (i % 2)] = -t if i % 2 == 0 else t
        for i, t in enumerate(en):
            peel[r, 4 * (i // 2) + 2 + (i % 2)] = -t if i % 2 == 0 else t
        # the same slot order for the CUDA clean-up, so that ties sort identically
        for s in range(32):
            v = peel[r, s]
            v = -v if s % 2 == 0 else v
            mine_t[r, s], mine_f[r, s] = v, (-1 if (s % 4) < 2 else 1)
    rt, rf = ref_gpu.ref_tide(peel, 2000.0)
    h = ctypes.c_void_p()
    _lib.check(lib.drr_create(0, ctypes.byref(h)))
    _lib.check(lib.drr_mesh_clean_hits(h, _lib.ptr(mine_t), _lib.ptr(mine_f), n_rays, 32, 2000.0, _lib.MEM_HOST), h)
    lib.drr_destroy(h)
    assert np.array_equal(mine_t, rt) and np.array_equal(mine_f, rf)      # bit-exact with kernelTide


def _prims_for_oracle(meshes, all_materials):
    out = []
    for m in meshes:
        tw = (np.asarray(m.triangles, dtype=np.float64).reshape(-1, 3) @ m.world_from_ijk.data[:3, :3].T + m.world_from_ijk.data[:3, 3]).reshape(-1, 3, 3)
        out.append({"tris_world": tw, "mat": all_materials.index(m.material), "density": m.density, "additive": m.additive,
                    "subtractive": m.subtractive, "layer": m.layer})
    return out


def _scene():
    ct = phantoms.thorax_volume((48, 48, 40), (8.5, 8.5, 10.0), seed=4)
    sv, sf = phantoms.screw_mesh(rings_per_mm=0.7, segments=24)
    screw = Mesh(sv, sf, material="titanium")                              # additive only (config 4)
    phantoms.place_kwire(screw, (-20.0, -60.0, 10.0), (0.2, 1.0, 0.1))     # same pose helper: tip + axis
    bv, bf = phantoms.icosphere(35.0, 2)
    ball = Mesh(bv, bf, material="lung", density=0.3, subtractive=True, layer=1)   # carve the CT, fill with lung
    ball.translate((40.0, 10.0, -5.0))
    cv, cf = phantoms.box_mesh((12.0, 30.0, 12.0))
    cavity = Mesh(cv, cf, material="bone", density=0.0, subtractive=True, layer=1)  # pure carving, overlaps the ball
    cavity.translate((55.0, 10.0, -5.0))
    iv, if_ = phantoms.icosphere(20.0, 2)
    inner = Mesh(iv, if_, material="bone", layer=0)                         # additive, partly inside the layer-1 carve
    inner.translate((25.0, 10.0, -5.0))
    return ct, [screw, ball, cavity, inner]


@pytest.mark.gpu
def test_ct_plus_meshes_matches_oracle():
    ct, meshes = _scene()
    carm = phantoms.MobileCArmGeometry(sensor_width=96, sensor_height=80, pixel_size=3.0)
    poses = phantoms.c2_poses(2, seed=8, carm=carm)
    W, H = 96, 80

    class Dev:
        source_to_detector_distance = carm.source_to_detector_distance
        camera_intrinsics = carm.camera_intrinsics
        detector_height, detector_width = carm.detector_height, carm.detector_width

    with Projector([ct] + meshes, device=Dev(), spectrum="90KV_AL40", neglog=False, step=0.25) as p:
        mats = p.all_materials
        area = p.project_line_integrals(*poses, max_ray_length=carm.max_ray_length)
        img = p.project(*poses, max_ray_length=carm.max_ray_length)
    assert mats == ["air", "bone", "lung", "soft tissue", "titanium"]
    st = cases.tables([ct], "90KV_AL40", None)
    from deepdrr_b200.material import absorb_coef_table
    from deepdrr_b200.scene import remap_labels

    mu = absorb_coef_table(mats, st.energies)
    labels = [remap_labels(ct, mats)]
    mesh_mats = sorted({mats.index(m.material) for m in meshes})
    checked = 0
    for n, pose in enumerate(poses):
        w2i, src, ijk = geo.pose_arrays(pose, [ct])
        mb = mesh_oracle.mesh_buffers(_prims_for_oracle(meshes, mats), w2i, pose.center_in_world, W, H, 2, 32,
                                      2 * carm.source_to_detector_distance, mesh_mats)
        r = cpu_oracle.project([ct.data], labels, len(mats), W, H, 0.25, w2i, src, ijk, carm.max_ray_length, st.energies, st.pdf, mu,
                               mesh=mb)
        # silhouette pixels: a hit count that differs between fp32 and fp64 edge tests shows up as a large difference
        diff = np.abs(area[n] - r.area) / np.maximum(np.abs(r.area), 1e-3)
        bad = diff.max(axis=0) > 1e-3
        assert bad.mean() < 0.02, f"too many silhouette mismatches: {bad.mean():.3f}"
        good = ~bad
        for m in range(len(mats)):
            mask = good & (r.area[m] > 0)
            if mask.any():
                rho_max = max([mm.density for mm in meshes if mm.material == mats[m]] + [1.0])
                atol = 8 * 6.1e-5 * rho_max / 10.0
                err = (np.abs(area[n, m].astype(np.float64) - r.area[m]) - atol)[mask] / r.area[m][mask]
                assert err.max() <= 1e-5, f"view {n} {mats[m]}: {err.max():.2e}"
                checked += int(mask.sum())
        # the absolute term above times titanium's mu/rho (several cm^2/g over most of the 90 kV spectrum)
        assert cases.rel_err(img[n], r.intensity)[good].max() <= 5e-4
        assert np.median(cases.rel_err(img[n], r.intensity)[good]) <= 1e-6
        ti = mats.index("titanium")
        assert (r.area[ti] > 0).sum() > 20 and (r.area[mats.index("lung")] > 0).sum() > 50    # the meshes are in view
    assert checked > 10000


@pytest.mark.gpu
def test_march_consumes_mesh_buffers_like_the_reference_kernel():
    """The same hit lists / additive buffers fed to the reference's projectKernel (MESH_ADDITIVE_ENABLED=1 cubin) and to
    the general march kernel through drr_set_mesh_buffers: line integrals must agree to 1e-5 everywhere."""
    import ctypes
    import torch
    from deepdrr_b200 import _lib
    from deepdrr_b200.scene import remap_labels
    from oracle import ref_gpu

    if not ref_gpu.available():
        pytest.skip("oracle/_ref not shipped")
    ct, meshes = _scene()
    carm = phantoms.MobileCArmGeometry(sensor_width=64, sensor_height=48, pixel_size=4.5)
    pose = phantoms.c2_poses(1, seed=8, carm=carm)[0]
    W, H = 64, 48
    mats = ["air", "bone", "lung", "soft tissue", "titanium"]
    labels = remap_labels(ct, mats)
    w2i, src, ijk = geo.pose_arrays(pose, [ct])
    mesh_mats = sorted({mats.index(m.material) for m in meshes})
    mb = mesh_oracle.mesh_buffers(_prims_for_oracle(meshes, mats), w2i, pose.center_in_world, W, H, 2, 32, 2 * carm.source_to_detector_distance, mesh_mats)
    ref = ref_gpu.RefProjector([ct.data], [labels], 5, lineint=True, variant="mesh")
    ref.set_mesh(mb, W * H)
    li = ref.line_integrals(W, H, 0.25, w2i, src, ijk, carm.max_ray_length)
    ref.close()
    # same buffers into the B200 library
    with Projector([ct], spectrum="90KV_AL40", neglog=False, step=0.25, camera_intrinsics=carm.camera_intrinsics) as p:
        p.all_materials = mats  # the material universe of the full scene
    with Projector([ct] + [mm for mm in meshes], spectrum="90KV_AL40", neglog=False, step=0.25, camera_intrinsics=carm.camera_intrinsics,
                   source_to_detector_distance=carm.source_to_detector_distance) as p:
        lib, h = _lib.load(), p._h
        _lib.check(lib.drr_set_meshes(h, 0, None, None, None, None, None, None, 2, 32), h)       # drop the library's own tracing
        p.meshes = []
        dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in mb.items()}
        _lib.check(lib.drr_set_mesh_buffers(h, 2, 32, dev["hit_alphas"].data_ptr(), dev["hit_facing"].data_ptr(), dev["layer_valid"].data_ptr(),
                                            dev["additive"].data_ptr(), dev["mesh_mats"].data_ptr(), len(mesh_mats), _lib.MEM_DEVICE), h)
        area = p.project_arrays(w2i[None], src[None], ijk[None], (W, H), carm.max_ray_length, want="area")[0]
    for m in range(5):
        mask = li[m] > 0
        assert np.all(area[m][~mask] == 0)
        if mask.any():
            assert cases.rel_err(area[m], li[m])[mask].max() <= 1e-5, mats[m]


@pytest.mark.gpu
def test_mesh_only_sphere_chord_and_enable_toggle():
    v, f = phantoms.icosphere(30.0, 3)
    ball = Mesh(v, f, material="iron", density=2.0)
    k = geo.CameraIntrinsicTransform.from_sizes((64, 64), 2.0, 1000.0)
    pose = phantoms.look_at_projection((0, -500.0, 0), (0, 1.0, 0), (0, 0, 1), k)
    with Projector([ball], camera_intrinsics=k, source_to_detector_distance=1000.0, neglog=False) as p:
        a = p.project_line_integrals(pose)[0, 0]
        ball.enabled = False
        z = p.project_line_integrals(pose)[0, 0]
    assert np.all(z == 0)
    # central ray: chord = 2R (a vertex-to-vertex diameter of the icosphere is exact) -> 2.0 g/cm^3 * 6 cm
    assert abs(a[32, 32] - 12.0) < 0.05 and abs(a[31, 31] - 12.0) < 0.05
    assert a[0, 0] == 0 and a.max() <= 12.0 + 1e-4
    # against the exact polyhedron (float64)
    w2i, _, _ = geo.pose_arrays(pose, [])
    t, ent = mesh_oracle.trace(ball.triangles, pose.center_in_world, mesh_oracle.pixel_dirs(w2i, 64, 64))
    s = np.where(ent, -1.0, 1.0)
    path = np.where(np.isfinite(t), t * s, 0.0).sum(axis=1).reshape(64, 64) * 2.0 / 10.0
    hits = np.isfinite(t).sum(axis=1).reshape(64, 64)
    ok = (hits == 2) & (np.abs(a - path) < 1e-3)
    assert ok.sum() > 0.95 * (hits == 2).sum()
    assert np.abs(a - path)[ok].max() <= 1e-5 * 12.0


def _query_scene():
    sv, sf = phantoms.screw_mesh(rings_per_mm=0.5, segments=16)
    screw = Mesh(sv, sf, material="titanium", tag="implant")
    phantoms.place_kwire(screw, (-20.0, -30.0, 10.0), (0.2, 1.0, 0.1))
    bv, bf = phantoms.icosphere(35.0, 2)
    ball = Mesh(bv, bf, material="lung", density=0.3, subtractive=True, additive=False, layer=1, tag="cavity")
    ball.translate((40.0, 10.0, -5.0))
    iv, if_ = phantoms.icosphere(20.0, 2)
    inner = Mesh(iv, if_, material="bone", layer=0, tag="implant")           # overlaps nothing of the screw
    inner.translate((25.0, 10.0, -5.0))
    ov, of = phantoms.box_mesh((10.0, 10.0, 10.0))
    other = Mesh(ov, of, material="iron", layer=0)                           # untagged
    other.translate((-40.0, 0.0, 30.0))
    return [screw, ball, inner, other]


@pytest.mark.gpu
def test_project_hits_travel_seg_match_restatement():
    """project_hits / project_travel / project_seg (reference projector.py:945-1053) per tag vs oracle.mesh_oracle.query."""
    meshes = _query_scene()
    W = H = 48
    k = geo.CameraIntrinsicTransform.from_sizes((W, H), 4.0, 1000.0)
    pose = phantoms.look_at_projection((30.0, -520.0, 10.0), (-0.03, 1.0, 0.0), (0, 0, 1), k)
    with Projector(meshes, camera_intrinsics=k, source_to_detector_distance=1000.0, neglog=False, max_mesh_hits=16) as p:
        tags = ["implant", "cavity", None, "nothing"]
        hits = p.project_hits(pose, tags=tags)
        travel = p.project_travel(pose, tags=tags)
        seg = p.project_seg(pose, tags=["implant", "cavity", "nothing"])
        assert p.project_hits(pose) == [] and p.project_travel(pose) == []
        with pytest.raises(NotImplementedError):
            p.project_hits(pose, pose, tags=tags)
        all_mats = p.all_materials
    prims = _prims_for_oracle(meshes, all_mats)
    w2i, _, _ = geo.pose_arrays(pose, [])
    src = pose.center_in_world
    for i, tag in enumerate(tags):
        sel = [tag is None or m.tag == tag for m in meshes]
        eh, cnt = mesh_oracle.query(prims, sel, "hits", w2i, src, W, H, 16, 2000.0)
        # silhouettes (fp32 vs fp64 edge tests disagree on the hit count) show up as a different number of finite slots
        same = np.isfinite(hits[i]).sum(axis=2) == np.isfinite(eh).sum(axis=2)
        assert same.mean() > 0.97
        a, b = hits[i][same], eh[same]
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin)
        assert np.all(np.abs(a[fin] - b[fin]) <= 8 * 6.1e-5)
        assert hits[i].dtype == np.float32 and hits[i].shape == (H, W, 16)
        et, cnt_t = mesh_oracle.query(prims, sel, "travel", w2i, src, W, H, 16, 2000.0)
        good = np.abs(travel[i] - et) <= 16 * 6.1e-5 + 1e-5 * et
        assert good.mean() > 0.97 and travel[i].min() >= 0
        if tag == "nothing":
            assert np.all(np.isinf(hits[i])) and np.all(travel[i] == 0)
        if tag == "cavity":                     # subtractive-only primitive: counted by hits, not drawn by the density pass
            assert np.all(travel[i] == 0) and np.isfinite(hits[i]).any()
    for i, tag in enumerate(["implant", "cavity", "nothing"]):
        sel = [m.tag == tag for m in meshes]
        es, cnt = mesh_oracle.query(prims, sel, "seg", w2i, src, W, H, 16, 2000.0)
        assert seg[i].dtype == np.uint8 and set(np.unique(seg[i])) <= {0, 255}
        assert (seg[i] == es).mean() > 0.985
        interior = cnt >= 2
        assert np.all(seg[i][interior & (es == 255)] == 255) or (seg[i] == es)[interior].mean() > 0.995
    assert seg[2].max() == 0 and seg[0].max() == 255
    # implant travel: closed additive meshes -> hits interval lengths add up to the travel length
    ih = hits[0]
    with np.errstate(invalid="ignore"):
        length = np.where(np.isfinite(ih[..., 1::2]), ih[..., 1::2] - ih[..., 0::2], 0.0).sum(axis=2)
    both = (np.isfinite(ih).sum(axis=2) % 2 == 0) & (travel[0] > 0)
    assert np.abs(length - travel[0])[both].max() < 1e-2


def test_bounding_sphere_in_frustum():
    """meshes_bounding_sphere_in_frustum (reference projector.py:882-943): sphere vs the four side planes."""
    from deepdrr_b200 import _lib
    v, f = phantoms.icosphere(10.0, 1)
    k = geo.CameraIntrinsicTransform.from_sizes((64, 48), 2.0, 1000.0)   # detector 128 x 96 mm at 1000 mm
    pose = phantoms.look_at_projection((0, -500.0, 0), (0, 1.0, 0), (0, 0, 1), k)
    p = Projector.__new__(Projector)
    p.initialized, p.device = True, None
    p._source_to_detector_distance = 1000.0
    def at(x, y, z):
        m = Mesh(v, f, material="iron")
        m.translate((x, y, z))
        return m
    # at y = 0 the frustum half-widths are 500 * 64/1000 = 32 mm (u) and 24 mm (v)
    axes = np.linalg.inv(pose.extrinsic.data)[:3, :3]       # camera x, y, z axes in world
    cx, cy = axes[:, 0], axes[:, 1]
    inside, touching, outside_u, outside_v, behind_edge = at(0, 0, 0), at(*(cx * 41.0)), at(*(cx * 43.0)), at(*(cy * 35.0)), at(*(cy * 33.0))
    res = p.meshes_bounding_sphere_in_frustum([inside, touching, outside_u, outside_v, behind_edge], pose)
    assert res == [True, True, False, False, True]
    r = inside.get_loose_bounding_sphere[1]
    assert abs(r - 10.0) < 1e-4
