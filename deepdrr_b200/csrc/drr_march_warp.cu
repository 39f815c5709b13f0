// Warp-cooperative single-volume ray march (the hot path of BASELINE C1/C2), libdrr_b200, sm_100a.
//
// Replaces the march of the reference's `projectKernel` (project_kernel.cu:334-553, "K.cu:n") for
// NUM_VOLUMES == 1 without meshes / outside air.  Same arithmetic as the per-ray kernel in
// drr_march.cu; different execution shape, chosen from its ncu profile (profiles/r01_*):
// per-lane "cell changed" reloads fired on 62 % of warp steps with 8 of 32 lanes active and cost more
// issue slots than the interpolation itself.
//
// Here a warp owns an 8x4 pixel tile and walks its 32 rays in lock step over the global step index
// t (all rays start at minAlpha ~ ray_length, so equal t means a coherent sample front).  For every
// segment of <= SEG steps the warp
//   1. bounds the voxel cells its rays will touch (two FMAs per lane and axis + redux.sync min/max),
//   2. stages those cells once, cooperatively, into shared memory: the 32 B filter-coefficient
//      record (cp.async, two planes) and the 1 B label code of each cell,
//   3. marches the segment out of shared memory with no divergence.  Where the whole box carries one
//      label and every lane is inside its window, groups of 8 steps take 4 samples from the texture
//      unit (tex3D) and 4 from the FMA pipes (the unit's arithmetic on the staged records; its 1.8
//      fixed-point coordinate comes from one round-down FFMA per axis, the record index from two
//      PRMT and a dp4a), so both resources are busy together; elsewhere groups of 4 steps look up
//      their label codes and either add four fetched values or replay the steps one by one.
// The sampler mix is a function of the step index only, so results do not depend on scheduling.
#include <math_constants.h>

#include <type_traits>

#include "drr_device.cuh"

#define TILE_W 8
#define TILE_H 4
#ifndef SEG_ALU
#define SEG_ALU 32                     // steps per segment, ALU sampler
#endif
#ifndef SEG_TEX
#define SEG_TEX 64                     // steps per segment, TEX sampler (stages label codes only)
#endif
#ifndef MAXC
#define MAXC 96                        // cells staged per warp and segment (ALU sampler: 32 B record + 1 B code each); a multiple of 16.
#endif                                 // 96 keeps 24 warps per SM under 100 KB of shared memory, i.e. 128 KB of L1 for the texture and
                                       // record traffic: 15.5 ms per C2 view against 17.0 with 256 cells (64 ... 112 within 1 %)
#define WARP_SMEM (MAXC * 32 + MAXC)   // coefficient records + codes
#define MAXC_TEX WARP_SMEM             // the TEX sampler uses the whole buffer for codes
#ifndef MIN_BLOCKS
#define MIN_BLOCKS 3
#endif
#ifndef WARPS_PER_BLOCK
#define WARPS_PER_BLOCK 8
#endif
#ifndef SWB
#define SWB 5                          // single-volume kernel: warps per block ...
#endif
#ifndef SMB
#define SMB 4                          // ... and resident blocks per SM it is compiled for (register budget 65536 / (32 SWB SMB))
#endif
#define MULTI_MAXV 4                   // volumes the multi-volume variant handles
#ifndef FAST_ROTATE
#define FAST_ROTATE 1                  // fast segments rotate only the ray state they use
#endif
#ifndef RAYS_PER_LANE
#define RAYS_PER_LANE 2                // single-volume kernel: a warp walks an 8 x (4 * R) pixel tile, R = 1 or RAYS_PER_LANE per launch
#endif
// Slack of the staged box (MarchParams::slack_lo / slack_hi, voxels) and of the window tests (slack_alpha, mm): they cover the
// drift of the accumulated fp32 alpha against alpha + S * step (<= S / 2 ulps of alpha; times |d| in voxels) and, on the high
// side, the 1 / 512 by which a fixed-point coordinate rounds up.  The host derives them per launch from the scene's largest
// alpha and finest voxel pitch (drr_capi.cu: march_slack) -- 0.01 / 0.0125 voxel and 0.01 mm for C2 -- and routes scenes that
// would need more than a quarter voxel to the per-ray kernels.

// lane -> pixel inside an 8 x 4 block.  The texture unit filters the lanes of a fetch in groups of four (lanes 4i .. 4i+3, two
// data-pipe wavefronts per group and fetch at best).  With `quads` such a group is a 2 x 2 block of pixels instead of a 4 x 1 run:
// its rays share a cell more often and the fetches cost fewer wavefronts -- 1 to 2 % of the single-volume march at every ray
// spacing from 0.12 (C2, 14.15 -> 14.02 ms per view) to 0.3 voxel (tools/quads_sweep.py).  The multi-volume kernel loses 3 % with
// it on C3 (1.01 -> 1.05 ms) and keeps the row-major layout.
__device__ __forceinline__ int lane_u(int lane, bool quads) {
    return quads ? ((lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4)) : (lane & (TILE_W - 1));  // quads: lane bits 0, 2, 4
}
__device__ __forceinline__ int lane_v(int lane, bool quads) {
    return quads ? (((lane >> 1) & 1) | ((lane >> 2) & 2)) : (lane >> 3);                         // quads: lane bits 1, 3
}

// tile id -> view and this lane's pixel (8x4 tiles, row-major per view)
__device__ __forceinline__ void tile_pixel(const MarchParams& P, unsigned tile, int lane, int& view, int& udx, int& vdx) {
    const int tiles_x = (P.W + TILE_W - 1) / TILE_W, tiles_y = (P.H + TILE_H - 1) / TILE_H;
    const unsigned tiles_per_view = (unsigned)tiles_x * tiles_y;
    view = (int)(tile / tiles_per_view);
    const unsigned tv = tile - (unsigned)view * tiles_per_view;
    const int ty = tv / tiles_x, tx = tv - ty * tiles_x;
    udx = tx * TILE_W + lane_u(lane, false);
    vdx = ty * TILE_H + lane_v(lane, false);
}

// ---- running-total bookkeeping (same order of fp32 adds as K.cu:544-546) ------------------------
template <int NM>
__device__ __forceinline__ void w_checkin(float cur, int& live, float* acc) {
#pragma unroll
    for (int m = 0; m < NM; m++) acc[m] = (live == m) ? cur : acc[m];
    live = -1;
}
template <int NM>
__device__ __forceinline__ float w_checkout(int label, int& live, const float* acc) {
    float cur = 0.0f;
#pragma unroll
    for (int m = 0; m < NM; m++) cur = (label == m) ? acc[m] : cur;
    live = label;
    return cur;
}

// Generic sample straight from global memory: mixed-label / clamped cells and half-weighted ends.  Out of line on
// purpose: it is the rare path, and inlining its ~160 instructions at every call site bloats the segment loops.
template <int NM>
struct AccV {
    float v[NM];
};

template <int NM, bool USE_TEX>
__device__ __noinline__ AccV<NM> w_slow_sample_nl(const VolDev* __restrict__ volp, float x, float y, float z, float weight, AccV<NM> acc) {
    const VolDev& vol = *volp;
    float px = __fsub_rn(x, 1.0f), py = __fsub_rn(y, 1.0f), pz = __fsub_rn(z, 1.0f);  // K.cu:402-404
    float bx = floorf(px), by = floorf(py), bz = floorf(pz);
    int ci = min(max((int)bx + 2, 0), vol.ni), cj = min(max((int)by + 2, 0), vol.nj), ck = min(max((int)bz + 2, 0), vol.nk);
    uint2 lab8 = __ldg(vol.celll + ((unsigned)ck * (unsigned)(vol.nj + 1) + (unsigned)cj) * (unsigned)(vol.ni + 1) + (unsigned)ci);  // < 2^31 cells (drr_add_volume)
    float seg[NM];
#pragma unroll
    for (int m = 0; m < NM; m++) seg[m] = 0.0f;
    seg_weights<NM>(__fsub_rn(px, bx), __fsub_rn(py, by), __fsub_rn(pz, bz), lab8, seg);
    float cx = __fadd_rn(px, 0.5f), cy = __fadd_rn(py, 0.5f), cz = __fadd_rn(pz, 0.5f);  // K.cu:542
    // mixed-label samples go through the texture unit whenever the volume has a texture, also under the FMA-pipe sampler:
    // the emulated filter is 1 ulp off in 0.2 % of fetches, which shows on pixels whose total for a material is tiny
    float rho = (USE_TEX || vol.tex != 0) ? tex3D<float>(vol.tex, cx, cy, cz) : hw_trilinear_raw(vol, cx, cy, cz);
    float wr = __fmul_rn(weight, rho);
#pragma unroll
    for (int m = 0; m < NM; m++) acc.v[m] = __fmaf_rn(wr, seg[m], acc.v[m]);
    return acc;
}

// out-of-line call (cold paths with many call sites)
template <int NM, bool USE_TEX>
__device__ __forceinline__ void w_slow_sample_call(const VolDev& vol, float x, float y, float z, float weight, float* acc) {
    AccV<NM> a;
#pragma unroll
    for (int m = 0; m < NM; m++) a.v[m] = acc[m];
    a = w_slow_sample_nl<NM, USE_TEX>(&vol, x, y, z, weight, a);
#pragma unroll
    for (int m = 0; m < NM; m++) acc[m] = a.v[m];
}

// inlined (the general-segment loop, where the call overhead shows)
// `fetched` / `have_fetched`: the density already fetched for this sample at (x - 0.5, y - 0.5, z - 0.5) by the group front.
// For x, y, z >= 1 that is the reference's own coordinate: x - 1 and (x - 1) + 0.5 are both exact there, so it equals x - 0.5.
template <int NM, bool USE_TEX>
__device__ __forceinline__ void w_slow_sample(const VolDev& vol, float x, float y, float z, float weight, float* acc, float fetched = 0.0f,
                                              bool have_fetched = false) {
    float px = __fsub_rn(x, 1.0f), py = __fsub_rn(y, 1.0f), pz = __fsub_rn(z, 1.0f);  // K.cu:402-404
    float bx = floorf(px), by = floorf(py), bz = floorf(pz);
    int ci = min(max((int)bx + 2, 0), vol.ni), cj = min(max((int)by + 2, 0), vol.nj), ck = min(max((int)bz + 2, 0), vol.nk);
    uint2 lab8 = __ldg(vol.celll + ((unsigned)ck * (unsigned)(vol.nj + 1) + (unsigned)cj) * (unsigned)(vol.ni + 1) + (unsigned)ci);  // < 2^31 cells (drr_add_volume)
    float seg[NM];
#pragma unroll
    for (int m = 0; m < NM; m++) seg[m] = 0.0f;
    seg_weights<NM>(__fsub_rn(px, bx), __fsub_rn(py, by), __fsub_rn(pz, bz), lab8, seg);
    float cx = __fadd_rn(px, 0.5f), cy = __fadd_rn(py, 0.5f), cz = __fadd_rn(pz, 0.5f);  // K.cu:542
    // mixed-label samples go through the texture unit whenever the volume has a texture, also under the FMA-pipe sampler:
    // the emulated filter is 1 ulp off in 0.2 % of fetches, which shows on pixels whose total for a material is tiny
    float rho = fetched;
    if (!(have_fetched && x >= 1.0f && y >= 1.0f && z >= 1.0f))
        rho = (USE_TEX || vol.tex != 0) ? tex3D<float>(vol.tex, cx, cy, cz) : hw_trilinear_raw(vol, cx, cy, cz);
    float wr = __fmul_rn(weight, rho);
#pragma unroll
    for (int m = 0; m < NM; m++) acc[m] = __fmaf_rn(wr, seg[m], acc[m]);
}

// KTEX = how many samples of every group of 8 consecutive steps are fetched by the texture unit; the others are
// interpolated on the FMA pipes from the staged cell records.  8 = TEX only (no records staged), 0 = ALU only.
// R = rays per lane: a warp walks R * 32 rays (an 8 x 4R pixel tile) through the same staged boxes, so the per-segment
// work (bounding, staging, classification) is shared by R * 32 * S samples; the sample loops run once per ray.
// MULTI: the scene has more volumes; [olo, ohi] is the hull of this ray's windows in the OTHER volumes.  Segments that
// come near it are marched step by step with the reference's priority pick over all volumes (K.cu:458-547); the
// caller has established that the shared label cache never serves foreign labels on this tile (see march_multi_kernel).
template <int NM, int KTEX, bool MULTI, int R>
__device__ __forceinline__ void march_core(const VolDev& vol, const float step, const float sx, const float sy, const float sz,
                                           const float (&dx)[R], const float (&dy)[R], const float (&dz)[R], const float (&lo)[R],
                                           const float (&hi)[R], float (&alpha)[R], const int (&num_steps)[R], float4* s_coef, uint8_t* s_code,
                                           int lane, float (&acc)[R][NM], const float slack_lo, const float slack_hi, const float slack_a,
                                           const MarchParams* MP = nullptr, unsigned tile = 0, float olo = 1.0f, float ohi = -1.0f) {
    static_assert(!MULTI || R == 1, "the multi-volume march walks one ray per lane");
    constexpr bool USE_TEX = KTEX > 0;      // general / slow samples go through the texture unit when there is one
    constexpr bool STAGE_COEF = KTEX < 8;
    int ns_max = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int m = 0; m < NM; m++) acc[r][m] = 0.0f;
        ns_max = max(ns_max, num_steps[r]);
    }
    const int t_end = __reduce_max_sync(0xffffffffu, ns_max);
    if (t_end == 0) return;

    // ---- before the volume: replay the fp32 alpha accumulation only (K.cu:552) ----
    // Lower bound of the first in-range step of each ray; `drift` bounds how far the accumulated
    // alpha can be from minAlpha + t*step.
    int t = 0;
    {
        int n0 = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (num_steps[r] > 0) {
                float drift = (float)num_steps[r] * 0x1p-24f * fmaxf(MULTI ? fmaxf(hi[r], ohi) : hi[r], 1.0f);
                int n = max(0, (int)floorf(fminf(__fdiv_rn(__fsub_rn(__fsub_rn(lo[r], alpha[r]), drift), step), 2.0e9f)) - 2);
                n = min(n, num_steps[r]);
                if (MULTI) {
                    if (lo[r] > hi[r]) n = num_steps[r];  // this ray never enters the volume
                    if (olo <= ohi) n = min(n, max(0, (int)floorf(fminf(__fdiv_rn(__fsub_rn(__fsub_rn(olo, alpha[r]), drift), step), 2.0e9f)) - 2));
                }
                n0 = min(n0, n);
            }
        }
        const int n_skip = __reduce_min_sync(0xffffffffu, n0);
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (num_steps[r] > 0) {  // (rays that do not exist have nothing to keep in step)
                if (alpha[r] >= 0x1p-100f) alpha[r] = alpha_jump(alpha[r], step, n_skip);  // == n_skip times alpha += step
                else for (int i = 0; i < n_skip; i++) alpha[r] = __fadd_rn(alpha[r], step);
            }
        }
        t = n_skip;
    }

    float cur[R];
    int live[R];
#pragma unroll
    for (int r = 0; r < R; r++) { cur[r] = 0.0f; live[r] = -1; }
    const int nxm = vol.ni - 2, nym = vol.nj - 2, nzm = vol.nk - 2;  // max cell base
    const float sxm = sx - 1.0f, sym = sy - 1.0f, szm = sz - 1.0f;
    // The sample loops below exist ONCE in the code and always work on ray 0 of the lane; after each pass the rays of the lane
    // are rotated by one place, so R passes serve every ray and leave the order as it was.  (Unrolling the passes instead doubles
    // the code of the active path past the 32 KB instruction cache: 2.4 "no instruction" stall cycles per issue, 30 % slower.)
    float rdx[R], rdy[R], rdz[R], rlo[R], rhi[R];
    int rns[R];
#pragma unroll
    for (int r = 0; r < R; r++) { rdx[r] = dx[r]; rdy[r] = dy[r]; rdz[r] = dz[r]; rlo[r] = lo[r]; rhi[r] = hi[r]; rns[r] = num_steps[r]; }
    auto rotate = [&]() {
        if (R > 1) {
            auto rot = [&](auto& x) {
                auto x0 = x[0];
#pragma unroll
                for (int r = 0; r + 1 < R; r++) x[r] = x[r + 1];
                x[R - 1] = x0;
            };
            rot(rdx); rot(rdy); rot(rdz); rot(rlo); rot(rhi); rot(rns); rot(alpha); rot(cur); rot(live);
#pragma unroll
            for (int m = 0; m < NM; m++) {
                const float a0 = acc[0][m];
#pragma unroll
                for (int r = 0; r + 1 < R; r++) acc[r][m] = acc[r + 1][m];
                acc[R - 1][m] = a0;
            }
        }
    };
    // fast segments touch only the direction, alpha and the checked-out total of a ray (the per-material totals only when a lane
    // changes material: then through the pass index)
    auto rotate_fast = [&]() {
        static_assert(!FAST_ROTATE || R <= 2, "the pass index picks the totals of ray R - 1");
        if (R > 1) {
            auto rot = [&](auto& x) {
                auto x0 = x[0];
#pragma unroll
                for (int r = 0; r + 1 < R; r++) x[r] = x[r + 1];
                x[R - 1] = x0;
            };
            rot(rdx); rot(rdy); rot(rdz); rot(alpha); rot(cur); rot(live);
        }
    };

    while (t < t_end) {
        // the march goes on to the far end of the farthest volume (K.cu:321-334); past this volume's window nothing is added
        if (MULTI) {
            if (__all_sync(0xffffffffu, t >= rns[0] || (alpha[0] > rhi[0] + slack_a && (olo > ohi || alpha[0] > ohi + slack_a)))) break;
            const int S0 = min(STAGE_COEF ? SEG_ALU : SEG_TEX, t_end - t);
            const float a1 = __fmaf_rn((float)S0, step, alpha[0]);
            const bool near_other = (t < rns[0]) && (olo <= ohi) && !(a1 < olo - slack_a) && !(alpha[0] > ohi + slack_a);
            if (__any_sync(0xffffffffu, near_other)) {
                // ---- mixed segment: priority pick over every volume, sample by sample ----------------
                w_checkin<NM>(cur[0], live[0], acc[0]);
                const MarchParams& P = *MP;
                const int V = P.V;
                const int last = rns[0] - 1;
                int view, udx, vdx;  // recomputed from the tile id rather than kept in registers through the march
                tile_pixel(P, tile, lane, view, udx, vdx);
                udx = min(udx, P.W - 1); vdx = min(vdx, P.H - 1);
                const ViewDev* mvw = P.views + view;
                const Ray ry = make_ray(mvw->w2i, udx, vdx);
                float dxs[MULTI_MAXV], dys[MULTI_MAXV], dzs[MULTI_MAXV], los[MULTI_MAXV], his[MULTI_MAXV];
#pragma unroll
                for (int i = 0; i < MULTI_MAXV; i++) {
                    dxs[i] = dys[i] = dzs[i] = 0.0f; los[i] = 1.0f; his[i] = -1.0f;  // empty window: never picked
                    if (i >= V || P.enabled[i] == 0) continue;
                    ray_dir_ijk(ry, mvw->ijk[i], dxs[i], dys[i], dzs[i]);
                    float lo_i, hi_i;
                    if (slab_test(dxs[i], dys[i], dzs[i], mvw->src[i][0], mvw->src[i][1], mvw->src[i][2], P.vol[i].ni, P.vol[i].nj, P.vol[i].nk,
                                  P.max_ray_length, lo_i, hi_i)) { los[i] = lo_i; his[i] = hi_i; }
                }
                // subtractive meshes (K.cu:498-517): the depth counters are a function of alpha alone, so they are
                // rebuilt here from the head of each hit list and advanced step by step
                const bool carve = P.layer_valid != nullptr;
                int hidx[4] = {0, 0, 0, 0}, hdep[4] = {0, 0, 0, 0};
                const size_t npix = (size_t)P.W * P.H;
                const size_t hbase = (((size_t)view * P.mesh_layers) * npix + (size_t)vdx * P.W + udx) * P.max_hits;
                float al = alpha[0];
                for (int s = 0; s < S0; s++, t++) {
                    bool inside_mesh = false;
                    if (carve) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (j >= P.mesh_layers || P.layer_valid[j] == 0) continue;
                            const float* ha = P.hit_alphas + hbase + (size_t)j * npix * P.max_hits;
                            const int8_t* hf = P.hit_facing + hbase + (size_t)j * npix * P.max_hits;
                            while (hidx[j] < P.max_hits && hf[hidx[j]] != 0 && ha[hidx[j]] < al) { hdep[j] += hf[hidx[j]]; hidx[j] += 1; }
                            if (hdep[j] > 0) inside_mesh = true;
                        }
                    }
                    if (t < rns[0] && !inside_mesh) {
                        int curr_priority = 0x7fffffff, n_at = 0;  // K.cu:458-496; priorities are distinct here (drr_capi.cu)
#pragma unroll
                        for (int i = 0; i < MULTI_MAXV; i++) {
                            if (al < los[i] || al > his[i]) continue;
                            if (P.priority[i] < curr_priority) { curr_priority = P.priority[i]; n_at = 1; }
                            else if (P.priority[i] == curr_priority) n_at += 1;
                        }
                        if (n_at > 0) {
                            const float weight = __fmul_rn(__fdiv_rn(1.0f, (float)n_at), (t == 0 || t == last) ? 0.5f : 1.0f);  // K.cu:530-537
#pragma unroll
                            for (int i = 0; i < MULTI_MAXV; i++) {
                                if (al < los[i] || al > his[i] || P.priority[i] != curr_priority) continue;
                                const float x = __fmaf_rn(al, dxs[i], mvw->src[i][0]), y = __fmaf_rn(al, dys[i], mvw->src[i][1]),
                                            z = __fmaf_rn(al, dzs[i], mvw->src[i][2]);
                                w_slow_sample_call<NM, USE_TEX>(P.vol[i], x, y, z, weight, acc[0]);
                            }
                        }
                    }
                    al = __fadd_rn(al, step);  // K.cu:552
                }
                alpha[0] = al;
                continue;
            }
        }
        // ---- 1. bound the cells of this segment ------------------------------------------------
        int S = min(STAGE_COEF ? SEG_ALU : SEG_TEX, t_end - t);
        const int cap = STAGE_COEF ? MAXC : MAXC_TEX;
        int blx, bly, blz, nx, ny, nz;
        bool any;
        // The staging index is built from one byte per axis and byte weights (dp4a): with the coefficient records the cap of
        // MAXC <= 255 cells implies that; the code-only boxes of the TEX sampler (cap 33 * MAXC) have to ask for it.
        auto box_fits = [&](int bx, int by, int bz) {
            if (STAGE_COEF && MAXC <= 255) return bx * by * bz <= cap;
            return bx * by * bz <= cap && max(bx, max(by, bz)) <= 255 && (bx * by <= 255 || bz == 1);
        };
        for (;;) {
            int lx = 0x7fffffff, ly = 0x7fffffff, lz = 0x7fffffff, hx = (int)0x80000000, hy = (int)0x80000000, hz = (int)0x80000000;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float a1 = __fmaf_rn((float)S, step, alpha[r]);
                const bool part = (t < rns[r]) && !(a1 < rlo[r] - slack_a) && !(alpha[r] > rhi[r] + slack_a);
                const float p0x = __fmaf_rn(alpha[r], rdx[r], sxm), p1x = __fmaf_rn(a1, rdx[r], sxm);  // p = x - 1 to within an ulp: a bound, the slack covers it
                const float p0y = __fmaf_rn(alpha[r], rdy[r], sym), p1y = __fmaf_rn(a1, rdy[r], sym);
                const float p0z = __fmaf_rn(alpha[r], rdz[r], szm), p1z = __fmaf_rn(a1, rdz[r], szm);
                if (part) {
                    // slack: the drift of the accumulated alpha against a1 (both sides) and, on the high side, the 1/512 by which a
                    // sample's fixed-point coordinate can round up into the next cell (slack_lo / slack_hi: drr_capi.cu, march_slack)
                    lx = min(lx, (int)floorf(fminf(p0x, p1x) - slack_lo)); hx = max(hx, (int)floorf(fmaxf(p0x, p1x) + slack_hi));
                    ly = min(ly, (int)floorf(fminf(p0y, p1y) - slack_lo)); hy = max(hy, (int)floorf(fmaxf(p0y, p1y) + slack_hi));
                    lz = min(lz, (int)floorf(fminf(p0z, p1z) - slack_lo)); hz = max(hz, (int)floorf(fmaxf(p0z, p1z) + slack_hi));
                }
            }
            // clamping commutes with the warp-wide min / max: do it once, on the reduced values
            blx = __reduce_min_sync(0xffffffffu, lx); bly = __reduce_min_sync(0xffffffffu, ly); blz = __reduce_min_sync(0xffffffffu, lz);
            int bhx = __reduce_max_sync(0xffffffffu, hx), bhy = __reduce_max_sync(0xffffffffu, hy), bhz = __reduce_max_sync(0xffffffffu, hz);
            any = bhx >= blx;
            if (!any) break;
            blx = max(-2, min(nxm, blx)); bly = max(-2, min(nym, bly)); blz = max(-2, min(nzm, blz));
            bhx = max(-2, min(nxm, bhx)); bhy = max(-2, min(nym, bhy)); bhz = max(-2, min(nzm, bhz));
            nx = bhx - blx + 1; ny = bhy - bly + 1; nz = bhz - blz + 1;
            if (box_fits(nx, ny, nz) || S == 1) break;
            S >>= 1;
        }
        if (!any) {  // nobody samples in this segment: just advance alpha
#pragma unroll
            for (int r = 0; r < R; r++)
                for (int s = 0; s < S; s++) alpha[r] = __fadd_rn(alpha[r], step);
            t += S;
            continue;
        }
        const int ncell = nx * ny * nz;
        if (!box_fits(nx, ny, nz)) {
            // Even a single step of this tile does not fit the staging buffer (rays far apart compared with
            // the voxel size): take the generic per-sample path for this step.  Correct for any geometry;
            // the host picks the per-ray kernel for such set-ups (drr_capi.cu: pick_variant).
#pragma unroll 1
            for (int pass = 0; pass < R; pass++) {
                w_checkin<NM>(cur[0], live[0], acc[0]);
                const int last = rns[0] - 1;
                int tt = t;
                for (int s = 0; s < S; s++, tt++) {
                    const bool inr = (tt < rns[0]) && !(alpha[0] < rlo[0]) && !(alpha[0] > rhi[0]);
                    if (inr) {
                        const float x = __fmaf_rn(alpha[0], rdx[0], sx), y = __fmaf_rn(alpha[0], rdy[0], sy), z = __fmaf_rn(alpha[0], rdz[0], sz);
                        w_slow_sample_call<NM, USE_TEX>(vol, x, y, z, (tt == 0 || tt == last) ? 0.5f : 1.0f, acc[0]);
                    }
                    alpha[0] = __fadd_rn(alpha[0], step);
                }
                rotate();
            }
            t += S;
            continue;
        }

        // ---- 2. stage the cells ----------------------------------------------------------------
        __syncwarp();
        int first_code = -1;
        bool same = true;
        {
            // e -> (cx, cy, cz) with multiply-shift divisions: m = trunc(65536 / n) + 2 overestimates 65536 / n by less than 3, so
            // (e * m) >> 16 == e / n as long as 3 * e * n < 65536 (e < 2112 cells for the TEX sampler's code-only boxes, n <= e)
            // ... which holds for e * n < 21845; larger boxes take the exact division
            const unsigned mnx = (unsigned)(65536.0f / (float)nx) + 2u, mny = (unsigned)(65536.0f / (float)ny) + 2u;
            const bool small_box = ncell * max(nx, ny) < 21845;
            const unsigned row_stride = (unsigned)(vol.ni + 1), slice_stride = row_stride * (unsigned)(vol.nj + 1);
            const unsigned cell0 = (unsigned)(blz + 2) * slice_stride + (unsigned)(bly + 2) * row_stride + (unsigned)(blx + 2);
            auto cell_of = [&](int e) {
                unsigned row, cz;
                if (small_box) { row = ((unsigned)e * mnx) >> 16; cz = (row * mny) >> 16; }
                else { row = (unsigned)e / (unsigned)nx; cz = row / (unsigned)ny; }
                const unsigned cx = (unsigned)e - row * (unsigned)nx, cy = row - cz * (unsigned)ny;
                return cell0 + cz * slice_stride + cy * row_stride + cx;
            };
            if (STAGE_COEF) {
                // all loads of the segment are put in flight at once: the 32 B records go global -> shared with
                // cp.async (no registers, no per-iteration round trip), the label codes through registers
                int cc[(MAXC + 31) / 32];
#pragma unroll
                for (int k = 0; k < (MAXC + 31) / 32; k++) {
                    const int e = lane + 32 * k;
                    cc[k] = -1;
                    if (e < ncell) {
                        const unsigned cell = cell_of(e);
                        cc[k] = __ldg(vol.cellcode + cell);
                        // the two 16 B halves of a record go to separate planes: consecutive cells then fall into
                        // distinct 16 B bank groups for the LDS.128 of the FMA-pipe sampler
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_coef + e);
                        const float4* src = vol.cellc + 2 * cell;
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)(MAXC * 16)), "l"(src + 1));
                    }
                }
                asm volatile("cp.async.commit_group;");
#pragma unroll
                for (int k = 0; k < (MAXC + 31) / 32; k++) {
                    if (cc[k] >= 0) {
                        s_code[lane + 32 * k] = (uint8_t)cc[k];
                        same = same && (first_code < 0 || cc[k] == first_code);
                        first_code = cc[k];
                    }
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            } else {
                for (int e = lane; e < ncell; e += 32) {
                    const int cc = __ldg(vol.cellcode + cell_of(e));
                    s_code[e] = (uint8_t)cc;
                    same = same && (first_code < 0 || cc == first_code);
                    first_code = cc;
                }
            }
        }
        __syncwarp();
        // one label for the whole box?  (lanes without a cell inherit lane 0's code)
        const int code0 = __shfl_sync(0xffffffffu, first_code, 0);
        if (first_code < 0) first_code = code0;
        const bool seg_uniform = __all_sync(0xffffffffu, same && first_code == code0) && code0 != 0xFF;
        // every ray inside its [lo, hi] window and away from its half-weighted end samples for the whole segment?
        bool lane_allin = t > 0;
#pragma unroll
        for (int r = 0; r < R; r++)
            lane_allin = lane_allin && (t + S < rns[r]) && !(alpha[r] < rlo[r]) && !(__fmaf_rn((float)S, step, alpha[r]) + slack_a > rhi[r]);
        const bool seg_allin = __all_sync(0xffffffffu, lane_allin);

        const float b1x = (float)(blx + 1), b1y = (float)(bly + 1), b1z = (float)(blz + 1);
        const unsigned cell_w = 1u | ((unsigned)min(nx, 255) << 8) | ((unsigned)min(nx * ny, 255) << 16);  // idx = cx + nx * cy + nx * ny * cz as one dp4a
        const float kfx = __fmaf_rn(-256.0f, b1x, 8388608.0f), kfy = __fmaf_rn(-256.0f, b1y, 8388608.0f), kfz = __fmaf_rn(-256.0f, b1z, 8388608.0f);
        if (seg_uniform && seg_allin) {
            // ---- 3a. fast segment: one label, no range checks ------------------------------------
            // The unit's 1.8 fixed-point coordinate relative to the box, Q = floor(256 * (x - b1) + 0.5), in ONE FFMA per axis:
            // kq = 2^23 + 0.5 - 256 * b1 is exact (b1 >= 2 on interior cells, so kq < 2^23 where the grid is 0.5), the FMA
            // adds it to 256 * x without intermediate rounding, the sum is >= 2^23 (grid 1) and rounding DOWN leaves
            // 2^23 + Q in the mantissa: low byte = fraction, next byte = cell (box sides are < 256 cells).
            const float kqx = __fadd_rn(kfx, 0.5f), kqy = __fadd_rn(kfy, 0.5f), kqz = __fadd_rn(kfz, 0.5f);
            const int t_stop = t + S;
#pragma unroll 1
            for (int pass = 0; pass < R; pass++) {
#if FAST_ROTATE
                if (live[0] != code0) {
                    if (R == 1 || pass == 0) {
                        w_checkin<NM>(cur[0], live[0], acc[0]);
                        cur[0] = w_checkout<NM>(code0, live[0], acc[0]);
                    } else {
                        w_checkin<NM>(cur[0], live[0], acc[R - 1]);
                        cur[0] = w_checkout<NM>(code0, live[0], acc[R - 1]);
                    }
                }
#else
                if (live[0] != code0) {
                    w_checkin<NM>(cur[0], live[0], acc[0]);
                    cur[0] = w_checkout<NM>(code0, live[0], acc[0]);
                }
#endif
                const float dxr = rdx[0], dyr = rdy[0], dzr = rdz[0];
                float al = alpha[0], c = cur[0];
                int tt = t;
                // (Folding the - 0.5 into the FMA's addend would save three FADDs per fetch and is the same number except where the
                // coordinate lies in [2^k, 2^k + 0.5) -- 0.5 % of the samples, half an ulp.  Measured: that moves enough fixed-point
                // coordinates on rays that run along such a band to take C2's worst pixel from 4e-7 to 6.4e-6 of its line integral.)
                auto tex_sample = [&](float a) {
                    const float x = __fmaf_rn(a, dxr, sx), y = __fmaf_rn(a, dyr, sy), z = __fmaf_rn(a, dzr, sz);
                    return tex3D<float>(vol.tex, __fsub_rn(x, 0.5f), __fsub_rn(y, 0.5f), __fsub_rn(z, 0.5f));  // K.cu:542
                };
                auto alu_sample = [&](float a) {
                    const float x = __fmaf_rn(a, dxr, sx), y = __fmaf_rn(a, dyr, sy), z = __fmaf_rn(a, dzr, sz);
                    const unsigned qx = __float_as_uint(__fmaf_rd(x, 256.0f, kqx)), qy = __float_as_uint(__fmaf_rd(y, 256.0f, kqy)),
                                   qz = __float_as_uint(__fmaf_rd(z, 256.0f, kqz));
                    const unsigned cells = __byte_perm(__byte_perm(qx, qy, 0x0051), qz, 0x0510);  // (cx, cy, cz, -)
                    const int idx = (int)__dp4a(cells, cell_w, 0u);  // inside the staged box: the host sizes the slack for that (march_slack)
                    const float4 cA = s_coef[idx], cB = s_coef[MAXC + idx];
                    return hw_trilinear_cell2q(qx, qy, qz, cA, cB);
                };
                if (KTEX == 8) {
#pragma unroll 4
                    for (; tt < t_stop; tt++) {
                        c = __fadd_rn(c, tex_sample(al));
                        al = __fadd_rn(al, step);  // K.cu:552
                    }
                } else if (KTEX == 0) {
#pragma unroll 2
                    for (; tt < t_stop; tt++) {
                        c = __fadd_rn(c, alu_sample(al));
                        al = __fadd_rn(al, step);
                    }
                } else {
                    // groups of 8 steps: the texture fetches are issued first, the FMA-pipe samples are computed while
                    // they are in flight, and the running total receives the eight values in step order (K.cu:544-546)
                    for (; tt + 8 <= t_stop; tt += 8) {
                        float a[8], v[8];
                        a[0] = al;
#pragma unroll
                        for (int j = 1; j < 8; j++) a[j] = __fadd_rn(a[j - 1], step);
                        // which steps of the group use the texture unit: spread evenly (Bresenham)
#pragma unroll
                        for (int j = 0; j < 8; j++) if (((j + 1) * KTEX) / 8 != (j * KTEX) / 8) v[j] = tex_sample(a[j]);
#pragma unroll
                        for (int j = 0; j < 8; j++) if (((j + 1) * KTEX) / 8 == (j * KTEX) / 8) v[j] = alu_sample(a[j]);
#pragma unroll
                        for (int j = 0; j < 8; j++) c = __fadd_rn(c, v[j]);
                        al = __fadd_rn(a[7], step);
                    }
                    for (; tt < t_stop; tt++) {
                        c = __fadd_rn(c, tex_sample(al));
                        al = __fadd_rn(al, step);
                    }
                }
                alpha[0] = al; cur[0] = c;
#if FAST_ROTATE
                rotate_fast();
#else
                rotate();
#endif
            }
            t = t_stop;
            continue;
        }

        // ---- 3b. general segment: per-sample label code and range check ----------------------------
        // four steps at a time, so that the texture fetches of a group are in flight together
#pragma unroll 1
        for (int pass = 0; pass < R; pass++) {
            const float dxr = rdx[0], dyr = rdy[0], dzr = rdz[0], lor = rlo[0], hir = rhi[0];
            const int nsr = rns[0], last = rns[0] - 1;
            float al = alpha[0], c = cur[0];
            int lv = live[0], tt = t;
            for (int s0 = 0; s0 < S; s0 += 4) {
                const int nb = min(4, S - s0);
                float aj[4], rj[4];
                aj[0] = al;
#pragma unroll
                for (int j = 1; j < 4; j++) aj[j] = __fadd_rn(aj[j - 1], step);  // K.cu:552
                // per step: in range?  (K.cu:472) -- then the texture fetches of the group, all in flight together -- then the label
                // code of every sample's cell.  Cell of p = x - 1 (K.cu:402-404, 416) inside the box: 2^23 + floor(256 * (x - b1))
                // from one round-down FFMA per axis (kf = 2^23 - 256 * b1, exact), cell index in the second byte.
                // Two copies of this front part: segments whose lanes are all inside their windows (most general segments are
                // "box not uniform" ones) need neither the range tests nor the end-sample test.
                bool inr[4];
                unsigned codes = 0;  // one byte per step
                bool plain = true;
                auto front = [&](auto all_inside) {
                    constexpr bool AI = decltype(all_inside)::value;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        inr[j] = AI ? true : (j < nb && (tt + j < nsr) && !(aj[j] < lor) && !(aj[j] > hir));
                        const float a = aj[j];
                        const float x = __fmaf_rn(a, dxr, sx), y = __fmaf_rn(a, dyr, sy), z = __fmaf_rn(a, dzr, sz);
                        rj[j] = 0.0f;
                        if (USE_TEX && inr[j]) rj[j] = tex3D<float>(vol.tex, __fsub_rn(x, 0.5f), __fsub_rn(y, 0.5f), __fsub_rn(z, 0.5f));  // K.cu:542
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float a = aj[j];
                        const float x = __fmaf_rn(a, dxr, sx), y = __fmaf_rn(a, dyr, sy), z = __fmaf_rn(a, dzr, sz);
                        const unsigned qx = __float_as_uint(__fmaf_rd(x, 256.0f, kfx)), qy = __float_as_uint(__fmaf_rd(y, 256.0f, kfy)),
                                       qz = __float_as_uint(__fmaf_rd(z, 256.0f, kfz));
                        int idx = (int)__dp4a(__byte_perm(__byte_perm(qx, qy, 0x0051), qz, 0x0510), cell_w, 0u);
                        idx = inr[j] ? min(idx, ncell - 1) : 0;
                        int code = s_code[idx];
                        if (!AI && ((tt + j == 0) | (tt + j == last))) code = 0xFF;  // half-weighted end samples take the generic path
                        codes |= (unsigned)code << (8 * j);
                        plain = plain && (!inr[j] || code == lv);
                    }
                };
                if (seg_allin && nb == 4) front(std::true_type{}); else front(std::false_type{});
                if (USE_TEX && __all_sync(0xffffffffu, plain)) {
                    // the whole warp stays on its materials for the group (lanes out of range fetched nothing: + 0)
#pragma unroll
                    for (int j = 0; j < 4; j++) c = __fadd_rn(c, rj[j]);
                    tt += nb;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (j < nb) {
                            const int code = (int)((codes >> (8 * j)) & 0xFFu);
                            if (inr[j]) {
                                if (code != lv) {
                                    w_checkin<NM>(c, lv, acc[0]);
                                    if (code != 0xFF) c = w_checkout<NM>(code, lv, acc[0]);
                                }
                                const float a = aj[j];
                                const float x = __fmaf_rn(a, dxr, sx), y = __fmaf_rn(a, dyr, sy), z = __fmaf_rn(a, dzr, sz);
                                if (code != 0xFF) {
                                    if (USE_TEX) {
                                        c = __fadd_rn(c, rj[j]);
                                    } else if (STAGE_COEF) {
                                        // fraction inside the sample's own cell: p - floor(p) with p = x - 1 is x - floor(x), exact; the cell's
                                        // place in the box from integers.  (x - b1 is NOT exact when the box starts in the clamped layers,
                                        // b1 <= 0: it once moved a fraction of 191.49994 / 256 to 191.5000013 / 256, one fixed-point step.)
                                        const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
                                        const float fbx = __fsub_rn(fx0, b1x), fby = __fsub_rn(fy0, b1y), fbz = __fsub_rn(fz0, b1z);
                                        const int idx = min((int)__fmaf_rn(__fmaf_rn(fbz, (float)ny, fby), (float)nx, fbx), MAXC - 1);
                                        const float4 cA = s_coef[idx], cB = s_coef[MAXC + idx];
                                        c = __fadd_rn(c, hw_trilinear_cell2(__fsub_rn(x, fx0), __fsub_rn(y, fy0), __fsub_rn(z, fz0), cA, cB));
                                    }
                                } else {
                                    w_slow_sample<NM, USE_TEX>(vol, x, y, z, (tt == 0 || tt == last) ? 0.5f : 1.0f, acc[0], rj[j], USE_TEX);  // K.cu:537
                                }
                            }
                            tt++;
                        }
                    }
                }
                float a_last = aj[0];
#pragma unroll
                for (int j = 1; j < 4; j++) a_last = (j < nb) ? aj[j] : a_last;
                al = __fadd_rn(a_last, step);
            }
            alpha[0] = al; cur[0] = c; live[0] = lv;
            rotate();
        }
        t += S;
    }
#pragma unroll
    for (int r = 0; r < R; r++) w_checkin<NM>(cur[r], live[r], acc[r]);
}

// Single-volume tile: per-ray set-up (K.cu:220-334), then the lock-step march.  Lane l walks pixels (udx, vdx + 4 r), r < R.
template <int NM, int KTEX, int R>
__device__ __forceinline__ void march_tile(const MarchParams& P, const ViewDev& vw, int udx, int vdx, float4* s_coef, uint8_t* s_code, int lane,
                                           float (&acc)[R][NM], unsigned long long& my_steps, unsigned long long& my_window) {
    const VolDev& vol = P.vol[0];
    float dx[R], dy[R], dz[R], lo[R], hi[R], alpha[R];
    int num_steps[R];
    const float sx = vw.src[0][0], sy = vw.src[0][1], sz = vw.src[0][2];
#pragma unroll
    for (int r = 0; r < R; r++) {
        dx[r] = dy[r] = dz[r] = lo[r] = alpha[r] = 0.f; hi[r] = -1.f; num_steps[r] = 0;
        const int v = vdx + TILE_H * r;
        if (udx < P.W && v < P.H && P.enabled[0] != 0) {
            Ray ry = make_ray(vw.w2i, udx, v);
            ray_dir_ijk(ry, vw.ijk[0], dx[r], dy[r], dz[r]);
            if (slab_test(dx[r], dy[r], dz[r], sx, sy, sz, vol.ni, vol.nj, vol.nk, P.max_ray_length, lo[r], hi[r])) {
                float minAlpha = fminf(ry.ray_length, lo[r]), maxAlpha = fmaxf(0.0f, hi[r]);   // K.cu:242-244, 321-322
                num_steps[r] = max((int)ceilf(__fdiv_rn(__fsub_rn(maxAlpha, minAlpha), P.step)), 0);  // K.cu:334
                alpha[r] = minAlpha;
            }
        }
        my_steps += (unsigned)num_steps[r];
        if (num_steps[r] > 0 && hi[r] >= lo[r])  // samples that fetch density: the steps whose alpha lies in [lo, hi] (to within one step)
            my_window += (unsigned)min(num_steps[r], (int)__fdiv_rn(__fsub_rn(hi[r], fmaxf(lo[r], alpha[r])), P.step) + 1);
    }
    march_core<NM, KTEX, false, R>(vol, P.step, sx, sy, sz, dx, dy, dz, lo, hi, alpha, num_steps, s_coef, s_code, lane, acc, P.slack_lo, P.slack_hi, P.slack_alpha);
}

// R rays per lane.  Two pay when neighbouring rays are close (C2: 0.12 voxel apart, the staged box hardly grows and its cost is
// shared by twice the rays); with rays a third of a voxel apart and more the 8 x 8 tile's boxes outgrow the staging buffer, the
// segments shrink, and one ray per lane is faster -- 2.2 against 3.2 ms for a 512^2 view of the C2 volume, 1.8 against 4.8 ms at
// 384^2 (gpurun_out/r2_rays_per_lane.log).  The host picks per batch (drr_capi.cu: MarchParams::rays_per_lane).
template <int NM, int R>
__global__ void __launch_bounds__(32 * SWB, SMB) march_warp_kernel(const __grid_constant__ MarchParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* s_coef = reinterpret_cast<float4*>(smem_raw + (size_t)warp * WARP_SMEM);
    uint8_t* s_code_alu = smem_raw + (size_t)warp * WARP_SMEM + MAXC * 32;
    uint8_t* s_code_tex = smem_raw + (size_t)warp * WARP_SMEM;
    const int tiles_x = (P.W + TILE_W - 1) / TILE_W, tiles_y = (P.H + TILE_H * R - 1) / (TILE_H * R);
    const unsigned tiles_per_view = (unsigned)tiles_x * tiles_y;
    const unsigned n_tiles = tiles_per_view * (unsigned)P.n_views;
    const size_t npix = (size_t)P.W * P.H;
    const float step = P.step;
    unsigned long long my_steps = 0, my_window = 0;
    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const unsigned view = tile / tiles_per_view;
        const unsigned tv = tile - view * tiles_per_view;
        const int ty = tv / tiles_x, tx = tv - ty * tiles_x;
        const int udx = tx * TILE_W + lane_u(lane, P.lane_quads != 0), vdx = ty * (TILE_H * R) + lane_v(lane, P.lane_quads != 0);
        // every tile uses the same sampler mix, and the mix is a function of the step index only, so results do not
        // depend on scheduling
        float acc[R][NM];
        const ViewDev& vw = P.views[view];
        switch (P.tex_eighths) {
            case 0: march_tile<NM, 0, R>(P, vw, udx, vdx, s_coef, s_code_alu, lane, acc, my_steps, my_window); break;
            case 1: case 2:
            case 3: march_tile<NM, 3, R>(P, vw, udx, vdx, s_coef, s_code_alu, lane, acc, my_steps, my_window); break;
            case 4: march_tile<NM, 4, R>(P, vw, udx, vdx, s_coef, s_code_alu, lane, acc, my_steps, my_window); break;
            case 5: case 6:
            case 7: march_tile<NM, 5, R>(P, vw, udx, vdx, s_coef, s_code_alu, lane, acc, my_steps, my_window); break;
            default: march_tile<NM, 8, R>(P, vw, udx, vdx, s_coef, s_code_tex, lane, acc, my_steps, my_window); break;
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int v = vdx + TILE_H * r;
            if (udx < P.W && v < P.H) {
                float* out = P.area + (size_t)view * P.M * npix + (size_t)v * P.W + udx;
#pragma unroll
                for (int m = 0; m < NM; m++) out[(size_t)m * npix] = __fdiv_rn(__fmul_rn(acc[r][m], step), 10.0f);  // K.cu:565-567, 582-584
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
        my_window += __shfl_xor_sync(0xffffffffu, my_window, o);
    }
    if (lane == 0 && my_steps) { atomicAdd(P.sample_count, my_steps); atomicAdd(P.sample_count + 1, my_window); }
}

#ifndef MARCH_WARP_R1_UNIT
// ---------------------------------------------------------------------------------------------
// Multi-volume scenes (BASELINE C3: a CT plus tool volumes).  Along most of every ray only one volume can be picked,
// and there the march is the single-volume one above -- provided the reference's shared label cache (K.cu:394-396,
// 416-431) never hands a volume another volume's labels.  With the volumes visited in index order every
// step, volume i can only start hitting the cache on foreign labels at a step where floor(p_i(t)) equals floor(p_j(t))
// for a j < i or floor(p_j(t-1)) for a j > i; both sides are straight lines in alpha, so "their difference stays
// outside the unit cube for the whole march" is a closed-form per-ray test.  A tile that passes it marches its longest
// volume in lock step and switches to per-sample priority picking (march_core, MULTI) wherever another volume's window
// is near; a tile that fails goes on a work list for march_general_list_kernel, which replays the reference step by step.
template <int NM>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK, MIN_BLOCKS) march_multi_kernel(const __grid_constant__ MarchParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* s_coef = reinterpret_cast<float4*>(smem_raw + (size_t)warp * WARP_SMEM);
    uint8_t* s_code_alu = smem_raw + (size_t)warp * WARP_SMEM + MAXC * 32;
    uint8_t* s_code_tex = smem_raw + (size_t)warp * WARP_SMEM;
    const int tiles_x = (P.W + TILE_W - 1) / TILE_W, tiles_y = (P.H + TILE_H - 1) / TILE_H;
    const unsigned tiles_per_view = (unsigned)tiles_x * tiles_y;
    const unsigned n_tiles = tiles_per_view * (unsigned)P.n_views;
    const size_t npix = (size_t)P.W * P.H;
    const float step = P.step;
    const int V = P.V;
    unsigned long long my_steps = 0;
    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const unsigned view = tile / tiles_per_view;
        const unsigned tv = tile - view * tiles_per_view;
        const int ty = tv / tiles_x, tx = tv - ty * tiles_x;
        const int udx = tx * TILE_W + lane_u(lane, false), vdx = ty * TILE_H + lane_v(lane, false);
        const bool ok = udx < P.W && vdx < P.H;
        const ViewDev& vw = P.views[view];

        // ---- every volume's slab test, as K.cu:265-334 ---------------------------------------------
        const Ray r = make_ray(vw.w2i, min(udx, P.W - 1), min(vdx, P.H - 1));
        float minAlpha = r.ray_length, maxAlpha = 0.0f;
        float dxs[MULTI_MAXV], dys[MULTI_MAXV], dzs[MULTI_MAXV], los[MULTI_MAXV], his[MULTI_MAXV];
        unsigned active = 0;  // bit i: this ray has a non-empty window in volume i
#pragma unroll
        for (int i = 0; i < MULTI_MAXV; i++) {
            dxs[i] = dys[i] = dzs[i] = 0.0f; los[i] = 1.0f; his[i] = -1.0f;
            if (i >= V || P.enabled[i] == 0) continue;
            ray_dir_ijk(r, vw.ijk[i], dxs[i], dys[i], dzs[i]);
            float lo_i, hi_i;
            if (slab_test(dxs[i], dys[i], dzs[i], vw.src[i][0], vw.src[i][1], vw.src[i][2], P.vol[i].ni, P.vol[i].nj, P.vol[i].nk,
                          P.max_ray_length, lo_i, hi_i)) {
                minAlpha = fminf(minAlpha, lo_i);
                maxAlpha = fmaxf(maxAlpha, hi_i);
                los[i] = lo_i; his[i] = hi_i;
                if (lo_i <= hi_i) active |= 1u << i;
            }
        }
        int num_steps = max((int)ceilf(__fdiv_rn(__fsub_rn(maxAlpha, minAlpha), step)), 0);
        if (!ok) { num_steps = 0; active = 0; }
        const unsigned tile_active = __reduce_or_sync(0xffffffffu, active);
        // Q3 test for every ordered pair (i contributes, j holds the cache): |p_i(alpha) - p_j(alpha - shift)| stays
        // outside the unit cube on [minAlpha, maxAlpha]
        bool suspect = false;
#pragma unroll
        for (int i = 0; i < MULTI_MAXV; i++) {
            if (!((active >> i) & 1u)) continue;
#pragma unroll
            for (int j = 0; j < MULTI_MAXV; j++) {
                if (j >= V || j == i) continue;
                const float shift = j > i ? step : 0.0f;  // volumes after i were last seen one step earlier
                const float di[3] = {dxs[i], dys[i], dzs[i]}, dj[3] = {dxs[j], dys[j], dzs[j]};
                float w_lo = minAlpha - 1.0f, w_hi = maxAlpha + 1.0f;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float A = (vw.src[i][c] - vw.src[j][c]) + shift * dj[c], B = di[c] - dj[c];
                    if (B != 0.0f) {
                        const float a0 = (-1.05f - A) / B, a1 = (1.05f - A) / B;
                        w_lo = fmaxf(w_lo, fminf(a0, a1));
                        w_hi = fminf(w_hi, fmaxf(a0, a1));
                    } else if (fabsf(A) >= 1.05f) {
                        w_hi = -INFINITY;
                    }
                }
                suspect = suspect || (w_lo <= w_hi);
            }
        }
        // where this ray can be inside a subtractive mesh (K.cu:498-517): from its first hit to its last one, or to the
        // end of the march if the entries and exits of a layer do not balance
        float mesh_lo = INFINITY, mesh_hi = -INFINITY;
        if (P.layer_valid != nullptr && ok) {
            const size_t pix = (size_t)vdx * P.W + udx;
            for (int j = 0; j < P.mesh_layers; j++) {
                if (P.layer_valid[j] == 0) continue;
                const float* ha = P.hit_alphas + (((size_t)view * P.mesh_layers + j) * npix + pix) * P.max_hits;
                const int8_t* hf = P.hit_facing + (((size_t)view * P.mesh_layers + j) * npix + pix) * P.max_hits;
                int depth = 0;
                for (int k = 0; k < P.max_hits && hf[k] != 0; k++) {
                    mesh_lo = fminf(mesh_lo, ha[k]); mesh_hi = fmaxf(mesh_hi, ha[k]);
                    depth += hf[k];
                }
                if (depth > 0) mesh_hi = INFINITY;
            }
        }
        if (__any_sync(0xffffffffu, suspect && num_steps > 0)) {
            if (lane == 0) P.worklist[atomicAdd(P.work_count, 1u)] = tile;
            continue;
        }
        float acc[NM];
        my_steps += (unsigned long long)num_steps * (unsigned)V;
        if (tile_active == 0) {
#pragma unroll
            for (int m = 0; m < NM; m++) acc[m] = 0.0f;
        } else {
            // the volume marched in lock step: the one with the longest window on this tile
            int a = 0;
            unsigned best = 0;
#pragma unroll
            for (int i = 0; i < MULTI_MAXV; i++) {
                const float len = ((active >> i) & 1u) ? fmaxf(his[i] - los[i], 0.0f) + 1.0f : 0.0f;
                const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(len));
                if (m > best) { best = m; a = i; }
            }
            float dx = 0.f, dy = 0.f, dz = 0.f, lo = 1.f, hi = -1.f, olo = INFINITY, ohi = -INFINITY;
#pragma unroll
            for (int i = 0; i < MULTI_MAXV; i++) {
                if (i == a) { dx = dxs[i]; dy = dys[i]; dz = dzs[i]; lo = los[i]; hi = his[i]; }
                else if ((active >> i) & 1u) { olo = fminf(olo, los[i]); ohi = fmaxf(ohi, his[i]); }
            }
            if (mesh_lo <= mesh_hi) { olo = fminf(olo, mesh_lo - 5.0f * P.slack_alpha); ohi = fmaxf(ohi, mesh_hi + 5.0f * P.slack_alpha); }
            const VolDev& vol = P.vol[a];
            const float sx = vw.src[a][0], sy = vw.src[a][1], sz = vw.src[a][2];
            const float dx1[1] = {dx}, dy1[1] = {dy}, dz1[1] = {dz}, lo1[1] = {lo}, hi1[1] = {hi};
            float al1[1] = {minAlpha};
            const int ns1[1] = {num_steps};
            float acc1[1][NM];
            if (P.tex_eighths <= 0)
                march_core<NM, 0, true, 1>(vol, step, sx, sy, sz, dx1, dy1, dz1, lo1, hi1, al1, ns1, s_coef, s_code_alu, lane, acc1, P.slack_lo, P.slack_hi, P.slack_alpha, &P, tile, olo, ohi);
            else if (P.tex_eighths >= 8)
                march_core<NM, 8, true, 1>(vol, step, sx, sy, sz, dx1, dy1, dz1, lo1, hi1, al1, ns1, s_coef, s_code_tex, lane, acc1, P.slack_lo, P.slack_hi, P.slack_alpha, &P, tile, olo, ohi);
            else
                march_core<NM, 5, true, 1>(vol, step, sx, sy, sz, dx1, dy1, dz1, lo1, hi1, al1, ns1, s_coef, s_code_alu, lane, acc1, P.slack_lo, P.slack_hi, P.slack_alpha, &P, tile, olo, ohi);
#pragma unroll
            for (int m = 0; m < NM; m++) acc[m] = acc1[0][m];
        }
        int view2, u2, v2;
        tile_pixel(P, tile, lane, view2, u2, v2);
        if (u2 < P.W && v2 < P.H) {
            const size_t pix = (size_t)v2 * P.W + u2;
            const int view = view2;
#pragma unroll
            for (int m = 0; m < NM; m++) acc[m] = __fmul_rn(acc[m], step);  // K.cu:565-567
            if (P.additive != nullptr) {                                    // K.cu:569-579
                const float* add = P.additive + (size_t)view * P.mesh_layers * P.n_mesh_mats * npix * 2;
                for (int i = 0; i < P.n_mesh_mats; i++)
                    for (int j = 0; j < P.mesh_layers; j++) {
                        const size_t idx = ((size_t)j * P.n_mesh_mats + i) * npix * 2 + pix * 2;
                        if (fabs((double)add[idx + 1]) < 0.00001) {
                            const int mm = P.mesh_mats[i];
                            const float v = fmaxf(add[idx], 0.0f);
#pragma unroll
                            for (int m = 0; m < NM; m++) if (m == mm) acc[m] = __fadd_rn(acc[m], v);
                        }
                    }
            }
            float* out = P.area + (size_t)view * P.M * npix + pix;
#pragma unroll
            for (int m = 0; m < NM; m++) out[(size_t)m * npix] = __fdiv_rn(acc[m], 10.0f);  // K.cu:582-584
        }
    }
    for (int o = 16; o > 0; o >>= 1) my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
    if (lane == 0 && my_steps) atomicAdd(P.sample_count, my_steps);
}

template <int NM>
static cudaError_t launch_multi_nm(const MarchParams& P, int n_sm, cudaStream_t s) {
    const size_t smem = (size_t)WARP_SMEM * WARPS_PER_BLOCK;
    cudaError_t e = cudaFuncSetAttribute(march_multi_kernel<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, march_multi_kernel<NM>, 32 * WARPS_PER_BLOCK, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    march_multi_kernel<NM><<<n_sm * occ, 32 * WARPS_PER_BLOCK, smem, s>>>(P);
    return cudaGetLastError();
}

cudaError_t drr_launch_march_multi(const MarchParams& P, int n_sm, cudaStream_t s) {
    if (P.V > MULTI_MAXV) return cudaErrorInvalidValue;
    switch (P.M) {
        case 1: return launch_multi_nm<1>(P, n_sm, s);
        case 2: return launch_multi_nm<2>(P, n_sm, s);
        case 3: return launch_multi_nm<3>(P, n_sm, s);
        case 4: return launch_multi_nm<4>(P, n_sm, s);
        case 5: return launch_multi_nm<5>(P, n_sm, s);
        case 6: return launch_multi_nm<6>(P, n_sm, s);
        case 7: return launch_multi_nm<7>(P, n_sm, s);
        case 8: return launch_multi_nm<8>(P, n_sm, s);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------------------------
#endif  // !MARCH_WARP_R1_UNIT
template <int NM, int R>
static cudaError_t launch_warp_nm_r(const MarchParams& P, int n_sm, cudaStream_t s) {
    const size_t smem = (size_t)WARP_SMEM * SWB;
    cudaError_t e = cudaFuncSetAttribute(march_warp_kernel<NM, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, march_warp_kernel<NM, R>, 32 * SWB, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    march_warp_kernel<NM, R><<<n_sm * occ, 32 * SWB, smem, s>>>(P);
    return cudaGetLastError();
}

// The one-ray-per-lane instantiations are compiled in their own translation unit (drr_march_warp_r1.cu includes this file with
// MARCH_WARP_R1_UNIT defined), so that the two halves build in parallel.
#ifdef MARCH_WARP_R1_UNIT
#define MARCH_WARP_R 1
#define MARCH_WARP_LAUNCH drr_launch_march_warp_r1
#else
#define MARCH_WARP_R RAYS_PER_LANE
#define MARCH_WARP_LAUNCH launch_march_warp_r2
cudaError_t drr_launch_march_warp_r1(const MarchParams& P, int n_sm, cudaStream_t s);
static
#endif
cudaError_t MARCH_WARP_LAUNCH(const MarchParams& P, int n_sm, cudaStream_t s) {
    switch (P.M) {
        case 1: return launch_warp_nm_r<1, MARCH_WARP_R>(P, n_sm, s);
        case 2: return launch_warp_nm_r<2, MARCH_WARP_R>(P, n_sm, s);
        case 3: return launch_warp_nm_r<3, MARCH_WARP_R>(P, n_sm, s);
        case 4: return launch_warp_nm_r<4, MARCH_WARP_R>(P, n_sm, s);
        case 5: return launch_warp_nm_r<5, MARCH_WARP_R>(P, n_sm, s);
        case 6: return launch_warp_nm_r<6, MARCH_WARP_R>(P, n_sm, s);
        case 7: return launch_warp_nm_r<7, MARCH_WARP_R>(P, n_sm, s);
        case 8: return launch_warp_nm_r<8, MARCH_WARP_R>(P, n_sm, s);
        default: return cudaErrorInvalidValue;
    }
}

#ifndef MARCH_WARP_R1_UNIT
cudaError_t drr_launch_march_warp(const MarchParams& P, int n_sm, cudaStream_t s) {
    return P.rays_per_lane == 1 ? drr_launch_march_warp_r1(P, n_sm, s) : launch_march_warp_r2(P, n_sm, s);
}
#endif  // !MARCH_WARP_R1_UNIT
