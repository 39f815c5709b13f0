// Mesh hit intervals by CUDA ray-triangle intersection (libdrr_b200, sm_100a).
//
// Replaces the reference's OpenGL path for meshes: pyrender / GLSL dual depth peeling
// (deepdrr/pyrenderdrr/renderer.py, shaders/*.frag), the GL<->CUDA interop copies (projector.py:71-113) and
// the three post-kernels of deepdrr/projector/peel_postprocess_kernel.cu ("PP.cu:n").  Semantics kept:
//   * rays and distances are the march's own: origin = source, unit direction from world_from_index,
//     distance = |hit - source| in world mm (shaders/density.frag:12);
//   * additive buffers per (layer, material): R = sum_hits d * s * rho, G = sum_hits s with s = +1 on
//     exit and -1 on entry (density.frag:11-12, blend ADD renderer.py:432-436), negative densities clamped
//     to 0 (renderer.py:424-425); projectKernel uses max(R, 0) iff |G| < 1e-5 (project_kernel.cu:569-579);
//   * subtractive layers: every hit of the layer's subtractive primitives, then exactly the clean-up of
//     `tide` (PP.cu:28-155): cut-off, selection sort, duplicate removal, compaction, altitude filter with
//     "sea level", depth > 1 removal, compaction; facing +1 = entry, -1 = exit (PP.cu:17-26);
//   * mesh-mesh subtraction (DRRMode.MESH_SUB, shaders/density_between.frag:29-51): a lower layer's
//     additive path does not count inside the cleaned intervals of higher subtractive layers.
// An exact enumerator has no pass limit, so where the rasteriser truncates to max_mesh_hits/4 peel passes
// this keeps the max_mesh_hits nearest hits instead (identical whenever the hits fit).
#include <math_constants.h>

#include "drr_device.cuh"

#define MESH_CHUNK 128
#define MESH_MAX_HITS 128

struct MeshPrimDev {
    int tri_begin, tri_end;
    int mat_slot;   // index into the mesh material list (additive buffers), -1 if none
    int layer;
    float density;
    int additive, subtractive;
};

// world = M(3x4) * local, one transform per (view, primitive)
__global__ void mesh_transform_kernel(const float* __restrict__ verts_local, const int* __restrict__ prim_of_tri,
                                      const float* __restrict__ world_from_mesh, int n_tris, int n_prims, int n_views,
                                      float* __restrict__ verts_world) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // vertex index
    const int view = blockIdx.y;
    if (i >= n_tris * 3) return;
    const int prim = prim_of_tri[i / 3];
    const float* M = world_from_mesh + ((size_t)view * n_prims + prim) * 12;
    const float x = verts_local[3 * i], y = verts_local[3 * i + 1], z = verts_local[3 * i + 2];
    float* o = verts_world + ((size_t)view * n_tris * 3 + i) * 3;
    o[0] = M[0] * x + M[1] * y + M[2] * z + M[3];
    o[1] = M[4] * x + M[5] * y + M[6] * z + M[7];
    o[2] = M[8] * x + M[9] * y + M[10] * z + M[11];
}

// Screen-space boxes: all rays leave one point, so a triangle can only be hit by the pixels its projection covers.
// tri_box[view][tri] = (u_min, u_max, v_min, v_max) in pixel indices with one pixel of slack; triangles with a vertex
// at or behind the source plane get an unbounded box.  prim_box[view][prim][4] is the union over the primitive.
#define BOX_INF (1 << 30)
__global__ void mesh_box_init_kernel(int* __restrict__ prim_box, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) prim_box[i] = (i & 1) ? -BOX_INF : BOX_INF;  // (min, max, min, max)
}

__global__ void mesh_project_kernel(const ViewDev* __restrict__ views, const float* __restrict__ source_world,
                                    const float* __restrict__ verts_world, const int* __restrict__ prim_of_tri, int n_tris, int n_prims,
                                    int4* __restrict__ tri_box, int* __restrict__ prim_box) {
    const int tri0 = blockIdx.x * blockDim.x + threadIdx.x, view = blockIdx.y;
    const bool live = tri0 < n_tris;
    const int tri = live ? tri0 : n_tris - 1;  // keep the warp whole for the reduction below
    const float* W = views[view].w2i_inv;
    const float* v = verts_world + ((size_t)view * n_tris + tri) * 9;
    const float ox = source_world[3 * view], oy = source_world[3 * view + 1], oz = source_world[3 * view + 2];
    float umin = 3e9f, umax = -3e9f, vmin = 3e9f, vmax = -3e9f;
    bool unbounded = false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float x = v[3 * k] - ox, y = v[3 * k + 1] - oy, z = v[3 * k + 2] - oz;
        const float qx = W[0] * x + W[1] * y + W[2] * z, qy = W[3] * x + W[4] * y + W[5] * z, qz = W[6] * x + W[7] * y + W[8] * z;
        if (!(qz > 1e-6f * (fabsf(qx) + fabsf(qy)) && qz > 0.0f)) { unbounded = true; continue; }
        const float pu = fminf(fmaxf(qx / qz, -1e9f), 1e9f), pv = fminf(fmaxf(qy / qz, -1e9f), 1e9f);
        umin = fminf(umin, pu); umax = fmaxf(umax, pu); vmin = fminf(vmin, pv); vmax = fmaxf(vmax, pv);
    }
    int4 b;
    if (unbounded) b = make_int4(-BOX_INF, BOX_INF, -BOX_INF, BOX_INF);
    else b = make_int4((int)floorf(umin - 1.5f), (int)ceilf(umax + 0.5f), (int)floorf(vmin - 1.5f), (int)ceilf(vmax + 0.5f));
    if (live) tri_box[(size_t)view * n_tris + tri] = b;
    // union over the primitive: one set of atomics per warp when the warp's triangles belong to one primitive
    const int prim = prim_of_tri[tri];
    int* pb = prim_box + ((size_t)view * n_prims + prim) * 4;
    if (__all_sync(0xffffffffu, prim == __shfl_sync(0xffffffffu, prim, 0))) {
        const int x0 = __reduce_min_sync(0xffffffffu, b.x), x1 = __reduce_max_sync(0xffffffffu, b.y);
        const int y0 = __reduce_min_sync(0xffffffffu, b.z), y1 = __reduce_max_sync(0xffffffffu, b.w);
        if ((threadIdx.x & 31) == 0) { atomicMin(pb + 0, x0); atomicMax(pb + 1, x1); atomicMin(pb + 2, y0); atomicMax(pb + 3, y1); }
    } else {
        atomicMin(pb + 0, b.x); atomicMax(pb + 1, b.y); atomicMin(pb + 2, b.z); atomicMax(pb + 3, b.w);
    }
}

// Stage those triangles of [base, base + cnt) whose box meets the tile [u0, u1] x [v0, v1], in index order (so sums
// over triangles keep their order).  All 128 threads of the block call it; returns the number staged.
__device__ __forceinline__ int stage_chunk(const float* __restrict__ vw, const int4* __restrict__ tb, int base, int cnt, int u0, int u1,
                                           int v0, int v1, float* s_v, int* s_cnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool keep = false;
    if (tid < cnt) {
        const int4 b = tb[base + tid];
        keep = b.x <= u1 && b.y >= u0 && b.z <= v1 && b.w >= v0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __syncthreads();  // the previous chunk has been consumed
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < warp; w++) off += s_cnt[w];
    const int total = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
    if (keep) {
        const float* src = vw + (size_t)(base + tid) * 9;
#pragma unroll
        for (int i = 0; i < 9; i++) s_v[off * 9 + i] = src[i];
    }
    __syncthreads();
    return total;
}

__device__ __forceinline__ bool prim_misses_tile(const int* __restrict__ pb, int u0, int u1, int v0, int v1) {
    return pb[0] > u1 || pb[1] < u0 || pb[2] > v1 || pb[3] < v0;
}

// Distance of a hit.  Moeller-Trumbore's own t = (e2 . q) / det takes q = s x e1 with |s| ~ 500 mm against sub-millimetre edges; on
// the sliver triangles of CAD meshes (the reference's 6.5 mm screw STL) that is 1.5e-2 mm off at the 90th percentile and 0.14 mm at
// worst in fp32.  A hit is rare (a few per ray against thousands of candidate triangles), so its distance is taken from the
// triangle's plane in double precision instead: t = n . (v0 - o) / n . d with n = e1 x e2 -- exact to the fp32 inputs (1e-5 mm),
// then rounded to the float the buffers hold.  Out of line: the fp64 code must not be if-converted into the per-candidate path.
__device__ __noinline__ float hit_distance(const float3& o, const float3& d, const float* __restrict__ v, float t_mt) {
    const double e1x = (double)v[3] - v[0], e1y = (double)v[4] - v[1], e1z = (double)v[5] - v[2];
    const double e2x = (double)v[6] - v[0], e2y = (double)v[7] - v[1], e2z = (double)v[8] - v[2];
    const double nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
    const double num = nx * ((double)v[0] - o.x) + ny * ((double)v[1] - o.y) + nz * ((double)v[2] - o.z);
    const double den = nx * d.x + ny * d.y + nz * d.z;
    return den != 0.0 ? (float)(num / den) : t_mt;
}

// Moeller-Trumbore, double sided.  Returns true and (t, entering) for t > 0.
__device__ __forceinline__ bool ray_tri(const float3& o, const float3& d, const float* __restrict__ v, float& t, bool& entering) {
    const float3 e1 = make_float3(v[3] - v[0], v[4] - v[1], v[5] - v[2]);
    const float3 e2 = make_float3(v[6] - v[0], v[7] - v[1], v[8] - v[2]);
    const float3 p = make_float3(d.y * e2.z - d.z * e2.y, d.z * e2.x - d.x * e2.z, d.x * e2.y - d.y * e2.x);
    const float det = e1.x * p.x + e1.y * p.y + e1.z * p.z;
    if (det == 0.0f) return false;
    const float inv = 1.0f / det;
    const float3 s = make_float3(o.x - v[0], o.y - v[1], o.z - v[2]);
    const float u = (s.x * p.x + s.y * p.y + s.z * p.z) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    const float3 q = make_float3(s.y * e1.z - s.z * e1.y, s.z * e1.x - s.x * e1.z, s.x * e1.y - s.y * e1.x);
    const float w = (d.x * q.x + d.y * q.y + d.z * q.z) * inv;
    if (w < 0.0f || u + w > 1.0f) return false;
    t = hit_distance(o, d, v, (e2.x * q.x + e2.y * q.y + e2.z * q.z) * inv);
    // geometric normal n = e1 x e2; det = d . (e1 x e2) ... sign(det) = sign(-d.n)?  p = d x e2, det = e1.(d x e2) = -d.(e1 x e2)
    entering = det > 0.0f;  // d . n < 0: the ray enters through an outward-facing (CCW) triangle
    return t > 0.0f;
}

// The clean-up of `tide` from the cut-off on (PP.cu:28-155), n = number of slots.
__device__ void tide_clean(float* ts, int8_t* facing, int n, float far_limit) {
    const float cutoffEpsilon = 0.00001f;
    for (int i = 0; i < n; i++)
        if (ts[i] < cutoffEpsilon || ts[i] > far_limit - 0.001f) { ts[i] = CUDART_INF_F; facing[i] = 0; }
    for (int sorted = 0; sorted < n; sorted++) {  // selection sort, same tie behaviour as the reference
        int minIdx = sorted;
        float minT = ts[minIdx];
        for (int i = sorted + 1; i < n; i++) { float t = ts[i]; if (t < minT) { minIdx = i; minT = t; } }
        float tmpT = ts[sorted]; ts[sorted] = minT; ts[minIdx] = tmpT;
        int8_t tf = facing[sorted]; facing[sorted] = facing[minIdx]; facing[minIdx] = tf;
    }
    {  // remove duplicates
        int dst = 0, src = 1;
        while (src < n) {
            if (ts[src] == ts[dst] && facing[src] == facing[dst]) { ts[src] = CUDART_INF_F; facing[src] = 0; src++; }
            else { dst = src; src++; }
        }
    }
    auto fill_gaps = [&]() {
        int dst = 0;
        while (dst < n && facing[dst] != 0) dst++;
        int src = dst + 1;
        while (src < n && dst < n) {
            while (src < n && facing[src] == 0) src++;
            if (src < n) { ts[dst] = ts[src]; facing[dst] = facing[src]; ts[src] = CUDART_INF_F; facing[src] = 0; }
            src++; dst++;
        }
    };
    fill_gaps();
    {
        int altitude = 0;
        for (int i = 0; i < n; i++) altitude += facing[i];
        const int seaLevel = max(0, altitude);
        int prev = 0, run = 0;
        for (int i = 0; i < n; i++) {
            run += facing[i];  // altitudes[i] of the reference, taken before this pass modifies facing[i]
            const int cur = run;
            if (cur < seaLevel || prev < seaLevel) { ts[i] = CUDART_INF_F; facing[i] = 0; }
            if (cur > 1 || prev > 1) { ts[i] = CUDART_INF_F; facing[i] = 0; }
            prev = cur;
        }
    }
    fill_gaps();
}

__global__ void tide_clean_kernel(float* __restrict__ ts, int8_t* __restrict__ facing, int n_rays, int n, float far_limit) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    float lt[MESH_MAX_HITS];
    int8_t lf[MESH_MAX_HITS];
    for (int i = 0; i < n; i++) { lt[i] = ts[(size_t)r * n + i]; lf[i] = facing[(size_t)r * n + i]; }
    tide_clean(lt, lf, n, far_limit);
    for (int i = 0; i < n; i++) { ts[(size_t)r * n + i] = lt[i]; facing[(size_t)r * n + i] = lf[i]; }
}

// Subtractive hit lists of one layer: block = 16x8 pixels, triangles streamed through shared memory.
__global__ void __launch_bounds__(128) mesh_subtractive_kernel(const ViewDev* __restrict__ views, const float* __restrict__ source_world,
                                                               const float* __restrict__ verts_world, const MeshPrimDev* __restrict__ prims,
                                                               int n_prims, int n_tris, int layer, int n_layers, int W, int H, int max_hits,
                                                               float far_limit, float* __restrict__ hit_alphas, int8_t* __restrict__ hit_facing,
                                                               const int4* __restrict__ tri_box, const int* __restrict__ prim_box) {
    __shared__ float s_v[MESH_CHUNK * 9];
    __shared__ int s_cnt[4];
    const int tiles_x = (W + 15) / 16;
    const int view = blockIdx.y;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int udx = tx * 16 + (threadIdx.x & 15), vdx = ty * 8 + (threadIdx.x >> 4);
    const bool ok = udx < W && vdx < H;
    Ray r = make_ray(views[view].w2i, min(udx, W - 1), min(vdx, H - 1));
    const float3 d = make_float3(r.rx, r.ry, r.rz);
    const float3 o = make_float3(source_world[3 * view], source_world[3 * view + 1], source_world[3 * view + 2]);
    float lt[MESH_MAX_HITS];
    int8_t lf[MESH_MAX_HITS];
    for (int i = 0; i < max_hits; i++) { lt[i] = CUDART_INF_F; lf[i] = 0; }
    int count = 0;
    const float* vw = verts_world + (size_t)view * n_tris * 9;
    for (int p = 0; p < n_prims; p++) {
        const MeshPrimDev pr = prims[p];
        if (!pr.subtractive || pr.layer != layer) continue;
        if (prim_misses_tile(prim_box + ((size_t)view * n_prims + p) * 4, tx * 16, tx * 16 + 15, ty * 8, ty * 8 + 7)) continue;
        for (int base = pr.tri_begin; base < pr.tri_end; base += MESH_CHUNK) {
            const int cnt = stage_chunk(vw, tri_box + (size_t)view * n_tris, base, min(MESH_CHUNK, pr.tri_end - base), tx * 16, tx * 16 + 15,
                                        ty * 8, ty * 8 + 7, s_v, s_cnt);
            if (!ok) continue;
            for (int k = 0; k < cnt; k++) {
                float t; bool entering;
                if (!ray_tri(o, d, s_v + 9 * k, t, entering)) continue;
                const int8_t f = entering ? 1 : -1;
                if (count < max_hits) { lt[count] = t; lf[count] = f; count++; }
                else {  // keep the max_hits nearest hits
                    int far_i = 0;
                    for (int i = 1; i < max_hits; i++) if (lt[i] > lt[far_i]) far_i = i;
                    if (t < lt[far_i]) { lt[far_i] = t; lf[far_i] = f; }
                }
            }
        }
    }
    if (!ok) return;
    tide_clean(lt, lf, max_hits, far_limit);
    const size_t npix = (size_t)W * H, pix = (size_t)vdx * W + udx;
    float* oa = hit_alphas + (((size_t)view * n_layers + layer) * npix + pix) * max_hits;
    int8_t* of = hit_facing + (((size_t)view * n_layers + layer) * npix + pix) * max_hits;
    for (int i = 0; i < max_hits; i++) { oa[i] = lt[i]; of[i] = lf[i]; }
}

// Additive buffers: R += g(t) * s * rho, G += s where g(t) = t minus the part of [0, t] covered by the
// cleaned intervals of higher subtractive layers (equivalent to density_between.frag for G = 0).
__global__ void __launch_bounds__(128) mesh_additive_kernel(const ViewDev* __restrict__ views, const float* __restrict__ source_world,
                                                            const float* __restrict__ verts_world, const MeshPrimDev* __restrict__ prims,
                                                            int n_prims, int n_tris, int n_layers, int n_mats, int W, int H, int max_hits,
                                                            const int8_t* __restrict__ layer_valid, const float* __restrict__ hit_alphas,
                                                            const int8_t* __restrict__ hit_facing, float* __restrict__ additive,
                                                            const int4* __restrict__ tri_box, const int* __restrict__ prim_box) {
    __shared__ float s_v[MESH_CHUNK * 9];
    __shared__ int s_cnt[4];
    const int tiles_x = (W + 15) / 16;
    const int view = blockIdx.y;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int udx = tx * 16 + (threadIdx.x & 15), vdx = ty * 8 + (threadIdx.x >> 4);
    const bool ok = udx < W && vdx < H;
    Ray r = make_ray(views[view].w2i, min(udx, W - 1), min(vdx, H - 1));
    const float3 d = make_float3(r.rx, r.ry, r.rz);
    const float3 o = make_float3(source_world[3 * view], source_world[3 * view + 1], source_world[3 * view + 2]);
    const size_t npix = (size_t)W * H, pix = (size_t)vdx * W + udx;
    const float* vw = verts_world + (size_t)view * n_tris * 9;
    for (int p = 0; p < n_prims; p++) {
        const MeshPrimDev pr = prims[p];
        if (!pr.additive || pr.mat_slot < 0 || pr.layer < 0 || pr.layer >= n_layers) continue;
        if (prim_misses_tile(prim_box + ((size_t)view * n_prims + p) * 4, tx * 16, tx * 16 + 15, ty * 8, ty * 8 + 7)) continue;
        const float rho = fmaxf(pr.density, 0.0f);  // renderer.py:424-425
        float R = 0.0f, G = 0.0f;
        for (int base = pr.tri_begin; base < pr.tri_end; base += MESH_CHUNK) {
            const int cnt = stage_chunk(vw, tri_box + (size_t)view * n_tris, base, min(MESH_CHUNK, pr.tri_end - base), tx * 16, tx * 16 + 15,
                                        ty * 8, ty * 8 + 7, s_v, s_cnt);
            if (!ok) continue;
            for (int k = 0; k < cnt; k++) {
                float t; bool entering;
                if (!ray_tri(o, d, s_v + 9 * k, t, entering)) continue;
                float g = t;
                for (int l = pr.layer + 1; l < n_layers; l++) {  // subtract the overlap with higher subtractive layers
                    if (!layer_valid[l]) continue;
                    const float* ha = hit_alphas + (((size_t)view * n_layers + l) * npix + pix) * max_hits;
                    const int8_t* hf = hit_facing + (((size_t)view * n_layers + l) * npix + pix) * max_hits;
                    for (int i = 0; i + 1 < max_hits; i += 2) {  // pairs (near, far) as kernelReorder2 hands them to GL
                        if (hf[i] == 0 || hf[i + 1] == 0) break;
                        const float nearD = ha[i], farD = ha[i + 1];
                        g -= fmaxf(0.0f, fminf(t, farD) - nearD);
                    }
                }
                const float s = entering ? -1.0f : 1.0f;  // density.frag: +1 on exit, -1 on entry
                R += g * s * rho;
                G += s;
            }
        }
        if (ok) {
            float* a = additive + ((((size_t)view * n_layers + pr.layer) * n_mats + pr.mat_slot) * npix + pix) * 2;
            a[0] += R;
            a[1] += G;
        }
    }
}

// Coverage mask (project_seg): 255 where any triangle of a selected primitive is hit, front or back face.
__global__ void __launch_bounds__(128) mesh_cover_kernel(const ViewDev* __restrict__ views, const float* __restrict__ source_world,
                                                         const float* __restrict__ verts_world, const MeshPrimDev* __restrict__ prims,
                                                         int n_prims, int W, int H, uint8_t* __restrict__ out,
                                                         const int4* __restrict__ tri_box, const int* __restrict__ prim_box) {
    __shared__ float s_v[MESH_CHUNK * 9];
    __shared__ int s_cnt[4];
    const int tiles_x = (W + 15) / 16;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int udx = tx * 16 + (threadIdx.x & 15), vdx = ty * 8 + (threadIdx.x >> 4);
    const bool ok = udx < W && vdx < H;
    Ray r = make_ray(views[0].w2i, min(udx, W - 1), min(vdx, H - 1));
    const float3 d = make_float3(r.rx, r.ry, r.rz);
    const float3 o = make_float3(source_world[0], source_world[1], source_world[2]);
    bool hit = false;
    for (int p = 0; p < n_prims; p++) {
        const MeshPrimDev pr = prims[p];
        if (!pr.subtractive) continue;  // the query marks the selected primitives this way
        if (prim_misses_tile(prim_box + (size_t)p * 4, tx * 16, tx * 16 + 15, ty * 8, ty * 8 + 7)) continue;
        for (int base = pr.tri_begin; base < pr.tri_end; base += MESH_CHUNK) {
            const int cnt = stage_chunk(verts_world, tri_box, base, min(MESH_CHUNK, pr.tri_end - base), tx * 16, tx * 16 + 15, ty * 8, ty * 8 + 7,
                                        s_v, s_cnt);
            if (!ok || hit) continue;
            for (int k = 0; k < cnt && !hit; k++) {
                float t; bool entering;
                hit = ray_tri(o, d, s_v + 9 * k, t, entering);
            }
        }
    }
    if (ok) out[(size_t)vdx * W + udx] = hit ? 255 : 0;
}

// project_travel epilogue (projector.py:1048-1051): drop unbalanced and negative pixels.
__global__ void mesh_travel_finish_kernel(const float* __restrict__ rg, int npix, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float v = rg[2 * i];
    if (fabsf(rg[2 * i + 1]) > 0.01f) v = 0.0f;
    if (v < 0.0f) v = 0.0f;
    out[i] = v;
}

// ---- launchers ---------------------------------------------------------------------------------
cudaError_t drr_launch_mesh_transform(const float* verts_local, const int* prim_of_tri, const float* world_from_mesh, int n_tris, int n_prims,
                                      int n_views, float* verts_world, cudaStream_t s) {
    dim3 grid((n_tris * 3 + 255) / 256, n_views);
    mesh_transform_kernel<<<grid, 256, 0, s>>>(verts_local, prim_of_tri, world_from_mesh, n_tris, n_prims, n_views, verts_world);
    return cudaGetLastError();
}

cudaError_t drr_launch_mesh_subtractive(const ViewDev* views, const float* source_world, const float* verts_world, const MeshPrimDev* prims,
                                        int n_prims, int n_tris, int layer, int n_layers, int W, int H, int n_views, int max_hits,
                                        float far_limit, float* hit_alphas, int8_t* hit_facing, const int4* tri_box, const int* prim_box,
                                        cudaStream_t s) {
    dim3 grid(((W + 15) / 16) * ((H + 7) / 8), n_views);
    mesh_subtractive_kernel<<<grid, 128, 0, s>>>(views, source_world, verts_world, prims, n_prims, n_tris, layer, n_layers, W, H, max_hits,
                                                 far_limit, hit_alphas, hit_facing, tri_box, prim_box);
    return cudaGetLastError();
}

cudaError_t drr_launch_mesh_additive(const ViewDev* views, const float* source_world, const float* verts_world, const MeshPrimDev* prims,
                                     int n_prims, int n_tris, int n_layers, int n_mats, int W, int H, int n_views, int max_hits,
                                     const int8_t* layer_valid, const float* hit_alphas, const int8_t* hit_facing, float* additive,
                                     const int4* tri_box, const int* prim_box, cudaStream_t s) {
    dim3 grid(((W + 15) / 16) * ((H + 7) / 8), n_views);
    mesh_additive_kernel<<<grid, 128, 0, s>>>(views, source_world, verts_world, prims, n_prims, n_tris, n_layers, n_mats, W, H, max_hits,
                                              layer_valid, hit_alphas, hit_facing, additive, tri_box, prim_box);
    return cudaGetLastError();
}

cudaError_t drr_launch_tide_clean(float* ts, int8_t* facing, int n_rays, int n, float far_limit, cudaStream_t s) {
    tide_clean_kernel<<<(n_rays + 63) / 64, 64, 0, s>>>(ts, facing, n_rays, n, far_limit);
    return cudaGetLastError();
}

cudaError_t drr_launch_mesh_cover(const ViewDev* views, const float* source_world, const float* verts_world, const MeshPrimDev* prims,
                                  int n_prims, int W, int H, uint8_t* out, const int4* tri_box, const int* prim_box, cudaStream_t s) {
    mesh_cover_kernel<<<((W + 15) / 16) * ((H + 7) / 8), 128, 0, s>>>(views, source_world, verts_world, prims, n_prims, W, H, out, tri_box,
                                                                       prim_box);
    return cudaGetLastError();
}

cudaError_t drr_launch_mesh_travel_finish(const float* rg, int npix, float* out, cudaStream_t s) {
    mesh_travel_finish_kernel<<<(npix + 255) / 256, 256, 0, s>>>(rg, npix, out);
    return cudaGetLastError();
}

// Fills tri_box [n_views][n_tris] and prim_box [n_views][n_prims][4] for the transformed vertices.
cudaError_t drr_launch_mesh_project(const ViewDev* views, const float* source_world, const float* verts_world, const int* prim_of_tri,
                                    int n_tris, int n_prims, int n_views, int4* tri_box, int* prim_box, cudaStream_t s) {
    const int n = n_views * n_prims * 4;
    mesh_box_init_kernel<<<(n + 255) / 256, 256, 0, s>>>(prim_box, n);
    if (n_tris > 0) {
        dim3 grid((n_tris + 255) / 256, n_views);
        mesh_project_kernel<<<grid, 256, 0, s>>>(views, source_world, verts_world, prim_of_tri, n_tris, n_prims, tri_box, prim_box);
    }
    return cudaGetLastError();
}
