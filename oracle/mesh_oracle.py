"""TEST INFRASTRUCTURE -- NumPy restatement of the reference's mesh semantics.

* ``tide_clean``      statement-by-statement restatement of ``tide`` from its cut-off step on
                      (deepdrr/projector/peel_postprocess_kernel.cu:28-155).
* ``trace``           brute-force float64 ray-triangle intersection: what the OpenGL rasteriser samples at
                      pixel centres (distance = |frag - cam|, shaders/density.frag:12).
* ``mesh_buffers``    the buffers ``projectKernel`` consumes: additive (R, G) per (layer, material)
                      (density.frag:11-12 + blend ADD), cleaned subtractive hit lists per layer, and the
                      mesh-mesh subtraction written literally as shaders/density_between.frag:29-51.
Never imported by the product.
"""
from __future__ import annotations

import numpy as np


def tide_clean(ts, facing, far_limit):
    ts = np.array(ts, dtype=np.float32)
    facing = np.array(facing, dtype=np.int8)
    n = len(ts)
    for i in range(n):  # PP.cu:28-36
        if ts[i] < np.float32(0.00001) or ts[i] > np.float32(far_limit) - np.float32(0.001):
            ts[i], facing[i] = np.inf, 0
    for s in range(n):  # PP.cu:39-60 selection sort
        mi, mt = s, ts[s]
        for i in range(s + 1, n):
            if ts[i] < mt:
                mi, mt = i, ts[i]
        ts[s], ts[mi] = mt, ts[s]
        facing[s], facing[mi] = facing[mi], facing[s]
    dst, src = 0, 1  # PP.cu:63-78
    while src < n:
        if ts[src] == ts[dst] and facing[src] == facing[dst]:
            ts[src], facing[src] = np.inf, 0
        else:
            dst = src
        src += 1

    def fill():  # PP.cu:80-104
        dst = 0
        while dst < n and facing[dst] != 0:
            dst += 1
        src = dst + 1
        while src < n and dst < n:
            while src < n and facing[src] == 0:
                src += 1
            if src < n:
                ts[dst], facing[dst] = ts[src], facing[src]
                ts[src], facing[src] = np.inf, 0
            src += 1
            dst += 1

    fill()
    alt = np.cumsum(facing.astype(np.int64))  # PP.cu:106-131
    sea = max(0, int(alt[-1]))
    prev = 0
    for i in range(n):
        cur = int(alt[i])
        if cur < sea or prev < sea:
            ts[i], facing[i] = np.inf, 0
        if cur > 1 or prev > 1:
            ts[i], facing[i] = np.inf, 0
        prev = cur
    fill()
    return ts, facing


def trace(tris, origin, dirs):
    """tris [m,3,3] world, origin [3], dirs [n,3] unit.  Returns (t [n,m] with inf for misses, entering [n,m])."""
    tris = np.asarray(tris, dtype=np.float64)
    o = np.asarray(origin, dtype=np.float64)
    d = np.asarray(dirs, dtype=np.float64)
    v0, e1, e2 = tris[:, 0], tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    p = np.cross(d[:, None, :], e2[None, :, :])
    det = np.einsum("mk,nmk->nm", e1, p)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / det
        s = o[None, :] - v0
        u = np.einsum("mk,nmk->nm", s, p) * inv
        q = np.cross(s, e1)
        w = np.einsum("nk,mk->nm", d, q) * inv
        t = np.einsum("mk,mk->m", e2, q)[None, :] * inv
    hit = (det != 0) & (u >= 0) & (u <= 1) & (w >= 0) & (u + w <= 1) & (t > 0)
    return np.where(hit, t, np.inf), det > 0


def pixel_dirs(w2i, W, H):
    u, v = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    m = np.asarray(w2i, dtype=np.float64).reshape(3, 3)
    r = np.stack([u, v, np.ones_like(u)], axis=-1) @ m.T
    return (r / np.linalg.norm(r, axis=-1, keepdims=True)).reshape(-1, 3)


def mesh_buffers(prims, w2i, source_world, W, H, n_layers, max_hits, far_limit, mesh_mats):
    """prims: list of dicts {tris_world [m,3,3], mat (global index), density, additive, subtractive, layer}.
    Returns dict(hit_alphas [L, H*W, max_hits], hit_facing, layer_valid [L], additive [L, n_mats, H*W, 2], mesh_mats)."""
    dirs = pixel_dirs(w2i, W, H)
    npix = W * H
    hit_alphas = np.zeros((n_layers, npix, max_hits), dtype=np.float32)
    hit_facing = np.zeros((n_layers, npix, max_hits), dtype=np.int8)
    layer_valid = np.zeros(n_layers, dtype=np.int8)
    additive = np.zeros((n_layers, len(mesh_mats), npix, 2), dtype=np.float64)
    traced = [trace(p["tris_world"], source_world, dirs) for p in prims]
    for l in range(n_layers):
        sub = [i for i, p in enumerate(prims) if p["subtractive"] and p["layer"] == l]
        if not sub:
            continue
        layer_valid[l] = 1
        for px in range(npix):
            ts, fs = [], []
            for i in sub:
                t, ent = traced[i]
                k = np.nonzero(np.isfinite(t[px]))[0]
                ts += list(t[px, k]); fs += [1 if e else -1 for e in ent[px, k]]
            order = np.argsort(ts, kind="stable")[:max_hits]
            lt = np.full(max_hits, np.inf, dtype=np.float32); lf = np.zeros(max_hits, dtype=np.int8)
            lt[:len(order)] = np.array(ts, dtype=np.float32)[order]; lf[:len(order)] = np.array(fs, dtype=np.int8)[order]
            hit_alphas[l, px], hit_facing[l, px] = tide_clean(lt, lf, far_limit)
    for i, p in enumerate(prims):
        if not p["additive"]:
            continue
        t, ent = traced[i]
        rho = max(p["density"], 0.0)
        s = np.where(ent, -1.0, 1.0)
        fin = np.isfinite(t)
        t = np.where(fin, t, 0.0)
        slot = list(mesh_mats).index(p["mat"])
        additive[p["layer"], slot, :, 0] += np.where(fin, t * s * rho, 0.0).sum(axis=1)   # density.frag
        additive[p["layer"], slot, :, 1] += np.where(fin, s, 0.0).sum(axis=1)
        for l in range(p["layer"] + 1, n_layers):                                          # density_between.frag, blend ADD
            if not layer_valid[l]:
                continue
            for j in range(0, max_hits - 1, 2):
                near, far = hit_alphas[l, :, j].astype(np.float64)[:, None], hit_alphas[l, :, j + 1].astype(np.float64)[:, None]
                valid = (hit_facing[l, :, j] != 0)[:, None] & (hit_facing[l, :, j + 1] != 0)[:, None] & fin
                near, far = np.where(np.isfinite(near), near, 0.0), np.where(np.isfinite(far), far, 0.0)
                res = np.where(t < near, -near * rho * s, 0.0) + np.where(t < far, far * rho * s, 0.0) + np.where((t >= near) & (t < far), -t * rho * s, 0.0)
                additive[p["layer"], slot, :, 0] += np.where(valid, res, 0.0).sum(axis=1)
    return {"hit_alphas": hit_alphas, "hit_facing": hit_facing, "layer_valid": layer_valid, "additive": additive.astype(np.float32),
            "mesh_mats": np.array(mesh_mats, dtype=np.int32)}


def query(prims, select, mode, w2i, source_world, W, H, max_hits, far_limit):
    """Restatement of Projector.project_hits / project_travel / project_seg for one tag (reference
    projector.py:945-1053; renderer.py:312-333 for which primitives a pass draws).
    prims as in mesh_buffers; select [n_prims] bool.  mode "hits" -> [H, W, max_hits] f32, "travel" -> [H, W] f32,
    "seg" -> [H, W] u8.  Also returns the per-pixel hit count of the drawn primitives (silhouette detection)."""
    dirs = pixel_dirs(w2i, W, H)
    npix = W * H
    if mode == "travel":
        drawn = [p for p, s in zip(prims, select) if s and p["additive"] and p["layer"] == 0]
    else:
        drawn = [p for p, s in zip(prims, select) if s]
    traced = [trace(p["tris_world"], source_world, dirs) for p in drawn]
    count = np.zeros(npix, dtype=np.int64)
    for t, _ in traced:
        count += np.isfinite(t).sum(axis=1)
    if mode == "seg":
        return np.where(count > 0, 255, 0).astype(np.uint8).reshape(H, W), count.reshape(H, W)
    if mode == "travel":
        R, G = np.zeros(npix), np.zeros(npix)
        for t, ent in traced:
            s = np.where(ent, -1.0, 1.0)
            fin = np.isfinite(t)
            R += np.where(fin, np.where(fin, t, 0.0) * s, 0.0).sum(axis=1)
            G += np.where(fin, s, 0.0).sum(axis=1)
        out = np.where(np.abs(G) > 0.01, 0.0, R)
        return np.maximum(out, 0.0).astype(np.float32).reshape(H, W), count.reshape(H, W)
    out = np.full((npix, max_hits), np.inf, dtype=np.float32)
    for px in range(npix):
        ts, fs = [], []
        for t, ent in traced:
            k = np.nonzero(np.isfinite(t[px]))[0]
            ts += list(t[px, k]); fs += [1 if e else -1 for e in ent[px, k]]
        if not ts:
            continue
        order = np.argsort(ts, kind="stable")[:max_hits]
        lt = np.full(max_hits, np.inf, dtype=np.float32); lf = np.zeros(max_hits, dtype=np.int8)
        lt[:len(order)] = np.array(ts, dtype=np.float32)[order]; lf[:len(order)] = np.array(fs, dtype=np.int8)[order]
        out[px], _ = tide_clean(lt, lf, far_limit)
    return out.reshape(H, W, max_hits), count.reshape(H, W)
