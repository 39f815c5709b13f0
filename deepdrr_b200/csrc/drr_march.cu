// Ray-march kernels of libdrr_b200 (sm_100a): replace the march of the reference's `projectKernel`
// (/root/reference/deepdrr/projector/project_kernel.cu:135-584, "K.cu:n").  Output: per-material area
// densities [view][M][H*W] in g/cm^2 (K.cu:565-584); the spectral kernel (drr_spectral.cu) takes over
// from there.
//
//   march_single_kernel<NM>   V == 1, no meshes, no outside-air term: the hot path (BASELINE C1/C2).
//       One ray per thread, persistent warps pulling 8x4-pixel tiles from an atomic queue.  The 8
//       texels + 8 labels of the current voxel cell live in registers (one 32 B + one 8 B record per
//       cell change), so the ~8-10 samples a ray spends in a cell touch no memory at all.  Warps are
//       split between two density samplers that compute the same arithmetic: the texture unit
//       (TEX role) and the SIMT emulation of the texture unit on cell records (ALU role), so both
//       the TEX pipe and the FMA pipes of every SM are busy.
//   march_general_kernel<NV, NM>  any volume count / priorities / meshes / outside air: exact
//       statement-by-statement semantics incl. the shared label cache.
#include <math_constants.h>

#include "drr_device.cuh"

#define TILE_W 8
#define TILE_H 4

// ---------------------------------------------------------------------------------------------
// hot path: single volume
// ---------------------------------------------------------------------------------------------
template <int NM>
struct CellState {
    float b1x, b1y, b1z;  // cell base + 1 (float), so that fr = x - b1 with x = p + 1 (p = K.cu:402 px)
    float4 c0, c1;        // filter coefficients (ALU role)
    uint2 lab8;           // 8 corner labels
    int label;            // the label when uniform
    bool slow;            // mixed labels or clamped (boundary) cell
};

template <int NM, bool USE_TEX>
__device__ __forceinline__ void reload_cell(const VolDev& vol, float x, float y, float z, CellState<NM>& cs) {
    // p = x - 1 (K.cu:402-404), base = floor(p) (K.cu:411-413)
    float bx = floorf(__fsub_rn(x, 1.0f)), by = floorf(__fsub_rn(y, 1.0f)), bz = floorf(__fsub_rn(z, 1.0f));
    cs.b1x = bx + 1.0f; cs.b1y = by + 1.0f; cs.b1z = bz + 1.0f;
    int ci = min(max((int)bx + 2, 0), vol.ni), cj = min(max((int)by + 2, 0), vol.nj), ck = min(max((int)bz + 2, 0), vol.nk);
    size_t cell = ((size_t)ck * (vol.nj + 1) + cj) * (vol.ni + 1) + ci;
    cs.lab8 = __ldg(vol.celll + cell);
    if (!USE_TEX) {
        const float4 A = __ldg(vol.cellc + 2 * cell), B = __ldg(vol.cellc + 2 * cell + 1);  // slices stored interleaved
        cs.c0 = make_float4(A.x, A.z, B.x, B.z);
        cs.c1 = make_float4(A.y, A.w, B.y, B.w);
    }
    unsigned l0 = cs.lab8.x & 0xFF;
    bool uniform = (cs.lab8.x == cs.lab8.y) && (cs.lab8.x == l0 * 0x01010101u);
    // clamped cells (base < 0) and the first voxel layer (base == 0, where the float rounding tricks of
    // hw_trilinear_cell lose their margin) take the integer path
    bool boundary = fminf(fminf(bx, by), bz) < 1.0f;
    cs.label = (int)l0;
    cs.slow = !uniform || (!USE_TEX && boundary);
}

// One generic sample (slow path: mixed-label or boundary cell, and the half-weighted end samples).
template <int NM, bool USE_TEX>
__device__ __forceinline__ void slow_sample(const VolDev& vol, float x, float y, float z, const CellState<NM>& cs, float weight,
                                            float* acc) {
    float px = __fsub_rn(x, 1.0f), py = __fsub_rn(y, 1.0f), pz = __fsub_rn(z, 1.0f);
    float bx = cs.b1x - 1.0f, by = cs.b1y - 1.0f, bz = cs.b1z - 1.0f;
    float seg[NM];
#pragma unroll
    for (int m = 0; m < NM; m++) seg[m] = 0.0f;
    seg_weights<NM>(__fsub_rn(px, bx), __fsub_rn(py, by), __fsub_rn(pz, bz), cs.lab8, seg);
    float cx = __fadd_rn(px, 0.5f), cy = __fadd_rn(py, 0.5f), cz = __fadd_rn(pz, 0.5f);  // K.cu:542
    // mixed-label samples go through the texture unit whenever the volume has a texture, also under the FMA-pipe sampler:
    // the emulated filter is 1 ulp off in 0.2 % of fetches, which shows on pixels whose total for a material is tiny
    float rho = (USE_TEX || vol.tex != 0) ? tex3D<float>(vol.tex, cx, cy, cz) : hw_trilinear_raw(vol, cx, cy, cz);
    float wr = __fmul_rn(weight, rho);
#pragma unroll
    for (int m = 0; m < NM; m++) acc[m] = __fmaf_rn(wr, seg[m], acc[m]);
}

// The running total of the current cell's material lives in one register (`cur`) while the ray is
// inside a uniform-label cell, so every sample is the same single fp32 add, in the same order, as
// the reference's `area_density[m] += ...` (K.cu:544-546).  live < 0: no material is checked out.
template <int NM>
__device__ __forceinline__ void checkin(float cur, int& live, float* acc) {
#pragma unroll
    for (int m = 0; m < NM; m++) acc[m] = (live == m) ? cur : acc[m];
    live = -1;
}
template <int NM>
__device__ __forceinline__ float checkout(int label, int& live, const float* acc) {
    float cur = 0.0f;
#pragma unroll
    for (int m = 0; m < NM; m++) cur = (label == m) ? acc[m] : cur;
    live = label;
    return cur;
}

template <int NM, bool USE_TEX>
__device__ __forceinline__ int march_ray_single(const MarchParams& P, const ViewDev& vw, int udx, int vdx, float* acc) {
    const VolDev& vol = P.vol[0];
    Ray r = make_ray(vw.w2i, udx, vdx);
#pragma unroll
    for (int m = 0; m < NM; m++) acc[m] = 0.0f;
    if (P.enabled[0] == 0) return 0;
    float dx, dy, dz;
    ray_dir_ijk(r, vw.ijk[0], dx, dy, dz);
    const float sx = vw.src[0][0], sy = vw.src[0][1], sz = vw.src[0][2];
    float lo, hi;
    if (!slab_test(dx, dy, dz, sx, sy, sz, vol.ni, vol.nj, vol.nk, P.max_ray_length, lo, hi)) return 0;
    const float minAlpha = fminf(r.ray_length, lo), maxAlpha = fmaxf(0.0f, hi);  // K.cu:242-244, 321-322
    const float step = P.step;
    const int num_steps = (int)ceilf(__fdiv_rn(__fsub_rn(maxAlpha, minAlpha), step));  // K.cu:334
    if (num_steps <= 0) return 0;

    CellState<NM> cs;
    cs.b1x = cs.b1y = cs.b1z = CUDART_INF_F;  // forces a reload at the first sample
    cs.slow = false; cs.label = 0; cs.lab8 = make_uint2(0, 0);
    cs.c0 = cs.c1 = make_float4(0, 0, 0, 0);
    float cur = 0.0f;
    int live = -1;
    float alpha = minAlpha;
    const int last = num_steps - 1;
    // Samples are only taken while lo <= alpha <= hi (K.cu:472).
    //  * The reference starts at minAlpha = min(ray_length, lo) (K.cu:242, 321); with the geometry
    //    library's world_from_index = R^T K^-1 the "ray_length" is ~1, so its march begins ~1 mm from
    //    the source and spends thousands of steps before the volume.  Those steps add nothing, but
    //    alpha is accumulated sequentially in fp32 (K.cu:552), so the accumulation is
    //    replayed here (one FADD per skipped step) to land on the same alpha grid.
    //  * alpha can also drift past hi during the last few steps: after n adds the drift is at most
    //    n * ulp(maxAlpha)/2, i.e. fewer than `margin` steps.  The main loop runs unchecked up to
    //    there; the tail is range-checked.
    const float drift = (float)num_steps * 0x1p-24f * fmaxf(maxAlpha, 1.0f);
    const int margin = (int)(__fdiv_rn(drift, step)) + 2;

    auto checked_sample = [&](int t) {
        float x = __fmaf_rn(alpha, dx, sx), y = __fmaf_rn(alpha, dy, sy), z = __fmaf_rn(alpha, dz, sz);
        float xr = __fsub_rn(x, cs.b1x), yr = __fsub_rn(y, cs.b1y), zr = __fsub_rn(z, cs.b1z);
        unsigned worst = max(max(__float_as_uint(xr), __float_as_uint(yr)), __float_as_uint(zr));
        if (!(alpha < lo) && !(alpha > hi)) {
            checkin<NM>(cur, live, acc);
            if (worst >= 0x3F800000u) reload_cell<NM, USE_TEX>(vol, x, y, z, cs);
            slow_sample<NM, USE_TEX>(vol, x, y, z, cs, (t == 0 || t == last) ? 0.5f : 1.0f, acc);  // K.cu:537
        }
        alpha = __fadd_rn(alpha, step);
    };

    int t = 0;
    while (t < num_steps && alpha < lo) {  // before the volume: replay the alpha accumulation only
        alpha = __fadd_rn(alpha, step);
        t++;
    }
    const int t_tail = max(t, num_steps - margin);
    if (t == 0 && t < t_tail) {  // first step of the march is half-weighted
        checked_sample(0);
        t = 1;
        cs.b1x = CUDART_INF_F;  // nothing is checked out: make the main loop reload its cell
    }
    for (; t < t_tail; t++) {
        float x = __fmaf_rn(alpha, dx, sx), y = __fmaf_rn(alpha, dy, sy), z = __fmaf_rn(alpha, dz, sz);
        float rho_tex = 0.0f;
        if (USE_TEX) rho_tex = tex3D<float>(vol.tex, __fsub_rn(x, 0.5f), __fsub_rn(y, 0.5f), __fsub_rn(z, 0.5f));
        float xr = __fsub_rn(x, cs.b1x), yr = __fsub_rn(y, cs.b1y), zr = __fsub_rn(z, cs.b1z);
        // inside the cell  <=>  all three in [0, 1): as unsigned bit patterns, [0,1) < 0x3F800000 and
        // negatives / NaN / >= 1 are above
        unsigned worst = max(max(__float_as_uint(xr), __float_as_uint(yr)), __float_as_uint(zr));
        if (worst >= 0x3F800000u) {
            checkin<NM>(cur, live, acc);
            reload_cell<NM, USE_TEX>(vol, x, y, z, cs);
            xr = __fsub_rn(x, cs.b1x); yr = __fsub_rn(y, cs.b1y); zr = __fsub_rn(z, cs.b1z);
            if (!cs.slow) cur = checkout<NM>(cs.label, live, acc);
        }
        if (cs.slow) {
            slow_sample<NM, USE_TEX>(vol, x, y, z, cs, 1.0f, acc);
        } else if (USE_TEX) {
            cur = __fadd_rn(cur, rho_tex);
        } else {
            cur = __fadd_rn(cur, hw_trilinear_cell(xr, yr, zr, cs.c0, cs.c1, 0.0f));
        }
        alpha = __fadd_rn(alpha, step);  // K.cu:552, sequential fp32
    }
    checkin<NM>(cur, live, acc);
    for (; t < num_steps; t++) checked_sample(t);
    return num_steps;
}

template <int NM>
__global__ void __launch_bounds__(256) march_single_kernel(const __grid_constant__ MarchParams P) {
    const int lane = threadIdx.x & 31;
    const int tiles_x = (P.W + TILE_W - 1) / TILE_W, tiles_y = (P.H + TILE_H - 1) / TILE_H;
    const unsigned tiles_per_view = (unsigned)tiles_x * tiles_y;
    const unsigned n_tiles = tiles_per_view * (unsigned)P.n_views;
    const size_t npix = (size_t)P.W * P.H;
    const float step = P.step;
    unsigned long long my_steps = 0;
    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const unsigned view = tile / tiles_per_view;
        const unsigned tv = tile - view * tiles_per_view;
        const int ty = tv / tiles_x, tx = tv - ty * tiles_x;
        const int udx = tx * TILE_W + (lane & (TILE_W - 1)), vdx = ty * TILE_H + (lane >> 3);
        const bool tex_role = ((tile * 5u) & 7u) < (unsigned)P.tex_eighths;  // per tile: results independent of scheduling
        if (udx < P.W && vdx < P.H) {
            float acc[NM];
            const ViewDev& vw = P.views[view];
            int ns = tex_role ? march_ray_single<NM, true>(P, vw, udx, vdx, acc) : march_ray_single<NM, false>(P, vw, udx, vdx, acc);
            my_steps += (unsigned)ns;
            float* out = P.area + (size_t)view * P.M * npix + (size_t)vdx * P.W + udx;
#pragma unroll
            for (int m = 0; m < NM; m++) out[(size_t)m * npix] = __fdiv_rn(__fmul_rn(acc[m], step), 10.0f);  // K.cu:565-567, 582-584
        }
    }
    // S_view bookkeeping: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
    if (lane == 0 && my_steps) atomicAdd(P.sample_count, my_steps);
}

// ---------------------------------------------------------------------------------------------
// general path: NV volumes, priorities, meshes, outside air
// ---------------------------------------------------------------------------------------------
// (NV, NM) are compile-time sizes of the per-ray register arrays.  Instantiated for every (V, M) up to 4 x 8 and once more "wide"
// as (DRR_MAX_VOLUMES, DRR_MAX_MATERIALS) for scenes beyond that, where the actual counts are read from P at run time -- the
// reference compiles its kernel for any -D NUM_VOLUMES / NUM_MATERIALS (projector.py:365-386).
template <int NV, int NM>
__device__ __forceinline__ void general_ray(const MarchParams& P, const int view, const int udx, const int vdx) {
    constexpr bool WIDE = NV > 4 || NM > 8;
    const ViewDev& vw = P.views[view];
    const size_t npix = (size_t)P.W * P.H;
    const size_t pix = (size_t)vdx * P.W + udx;
    const float step = P.step;

    Ray r = make_ray(vw.w2i, udx, vdx);
    float minAlpha = r.ray_length, maxAlpha = 0.0f;
    float area[NM];
#pragma unroll
    for (int m = 0; m < NM; m++) area[m] = 0.0f;
    float dx[NV], dy[NV], dz[NV], lo[NV], hi[NV];
    bool trace[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) {
        dx[i] = dy[i] = dz[i] = 0.0f; lo[i] = hi[i] = 0.0f;
        trace[i] = false;
        if (WIDE && i >= P.V) continue;  // the wide instantiation serves any V <= NV
        if (P.enabled[i] == 0) continue;
        ray_dir_ijk(r, vw.ijk[i], dx[i], dy[i], dz[i]);
        trace[i] = slab_test(dx[i], dy[i], dz[i], vw.src[i][0], vw.src[i][1], vw.src[i][2], P.vol[i].ni, P.vol[i].nj, P.vol[i].nk,
                             P.max_ray_length, lo[i], hi[i]);
        if (trace[i]) {
            minAlpha = fminf(minAlpha, lo[i]);
            maxAlpha = fmaxf(maxAlpha, hi[i]);
        }
    }
    const int num_steps = (int)ceilf(__fdiv_rn(__fsub_rn(maxAlpha, minAlpha), step));
    float alpha = minAlpha;
    if (P.attenuate_outside) {  // K.cu:359-361 (double arithmetic through the 0.1129 literal)
        int ai = P.air_index;
        float pre = __fdiv_rn(minAlpha, step);
#pragma unroll
        for (int m = 0; m < NM; m++) if (m == ai) area[m] = (float)((double)area[m] + (double)pre * 0.1129);
    }

    // shared label cache (K.cu:394-396, 416-431): integer corner coordinates of the last miss and
    // the volume they were fetched from
    int c_lo[3] = {0, 0, 0}, c_hi[3] = {0, 0, 0}, owner = -1;
    float prev[3] = {-1.0f, -1.0f, -1.0f};
    int hit_depth[4] = {0, 0, 0, 0}, hit_index[4] = {0, 0, 0, 0};
    const bool meshes = P.layer_valid != nullptr;
    const size_t view_hits = (size_t)view * P.mesh_layers * npix * P.max_hits;

    for (int t = 0; t < num_steps; t++) {
        // priority pick (K.cu:458-496); the "any_seg > 0" test is always true (labels are never null)
        int curr_priority = NV, n_at = 0;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            if (WIDE && i >= P.V) continue;
            if (!trace[i] || P.enabled[i] == 0) continue;
            if (alpha < lo[i] || alpha > hi[i]) continue;
            if (P.priority[i] < curr_priority) { curr_priority = P.priority[i]; n_at = 1; }
            else if (P.priority[i] == curr_priority) n_at += 1;
        }
        bool inside_mesh = false;  // K.cu:498-517
        if (meshes) {
            for (int j = 0; j < P.mesh_layers; j++) {
                if (P.layer_valid[j] == 0) continue;
                const float* ha = P.hit_alphas + view_hits + ((size_t)j * npix + pix) * P.max_hits;
                const int8_t* hf = P.hit_facing + view_hits + ((size_t)j * npix + pix) * P.max_hits;
                while (hit_index[j] < P.max_hits && hf[hit_index[j]] != 0 && ha[hit_index[j]] < alpha) {
                    hit_depth[j] += hf[hit_index[j]];
                    hit_index[j] += 1;
                }
                if (hit_depth[j] > 0) inside_mesh = true;
            }
        }
        float weight = 0.0f;
        if (n_at > 0) {
            weight = __fdiv_rn(1.0f, (float)n_at);
            weight = __fmul_rn(weight, (t == 0 || t == num_steps - 1) ? 0.5f : 1.0f);
        }
#pragma unroll
        for (int i = 0; i < NV; i++) {
            if (WIDE && i >= P.V) continue;  // (a volume that does not exist must not touch the shared label cache either)
            const VolDev& vol = P.vol[i];
            float px = __fsub_rn(__fmaf_rn(alpha, dx[i], vw.src[i][0]), 1.0f);
            float py = __fsub_rn(__fmaf_rn(alpha, dy[i], vw.src[i][1]), 1.0f);
            float pz = __fsub_rn(__fmaf_rn(alpha, dz[i], vw.src[i][2]), 1.0f);
            float bx = floorf(px), by = floorf(py), bz = floorf(pz);
            if (bx != prev[0] || by != prev[1] || bz != prev[2]) {
                prev[0] = bx; prev[1] = by; prev[2] = bz;
                owner = i;
                c_lo[0] = (int)bx; c_lo[1] = (int)by; c_lo[2] = (int)bz;
                c_hi[0] = (int)floorf(__fadd_rn(px, 1.0f));
                c_hi[1] = (int)floorf(__fadd_rn(py, 1.0f));
                c_hi[2] = (int)floorf(__fadd_rn(pz, 1.0f));
            }
            const bool contributes = !inside_mesh && n_at > 0 && trace[i] && P.priority[i] == curr_priority && P.enabled[i] == 1;
            if (!contributes) continue;
            // labels of the cache owner (normally this volume)
            uint2 lab8 = make_uint2(0, 0);
            {
                const VolDev& ov = P.vol[owner < 0 ? i : owner];
#pragma unroll
                for (int c = 0; c < 2; c++)
#pragma unroll
                    for (int b = 0; b < 2; b++)
#pragma unroll
                        for (int a = 0; a < 2; a++) {
                            unsigned l = (unsigned)label_at(ov, a ? c_hi[0] : c_lo[0], b ? c_hi[1] : c_lo[1], c ? c_hi[2] : c_lo[2]);
                            if (c) lab8.y |= l << (8 * (a + 2 * b)); else lab8.x |= l << (8 * (a + 2 * b));
                        }
            }
            float seg[NM];
#pragma unroll
            for (int m = 0; m < NM; m++) seg[m] = 0.0f;
            seg_weights<NM>(__fsub_rn(px, bx), __fsub_rn(py, by), __fsub_rn(pz, bz), lab8, seg);
            const float tcx = __fadd_rn(px, 0.5f), tcy = __fadd_rn(py, 0.5f), tcz = __fadd_rn(pz, 0.5f);  // K.cu:542
            float rho = vol.tex != 0 ? tex3D<float>(vol.tex, tcx, tcy, tcz) : hw_trilinear_raw(vol, tcx, tcy, tcz);
            float wr = __fmul_rn(weight, rho);
#pragma unroll
            for (int m = 0; m < NM; m++) area[m] = __fmaf_rn(wr, seg[m], area[m]);
        }
        if (!inside_mesh && n_at == 0 && P.attenuate_outside) {  // K.cu:522-527
            int ai = P.air_index;
#pragma unroll
            for (int m = 0; m < NM; m++) if (m == ai) area[m] = (float)((double)area[m] + 0.1129);
        }
        alpha = __fadd_rn(alpha, step);
    }
    if (P.attenuate_outside) {  // K.cu:558-560
        int ai = P.air_index;
        float extra = __fdiv_rn(__fsub_rn(r.ray_length, maxAlpha), step);
#pragma unroll
        for (int m = 0; m < NM; m++) if (m == ai) area[m] = (float)((double)area[m] + (double)extra * 0.1129);
    }
#pragma unroll
    for (int m = 0; m < NM; m++) area[m] = __fmul_rn(area[m], step);  // K.cu:565-567
    if (P.additive != nullptr) {                                      // K.cu:569-579
        const float* add = P.additive + (size_t)view * P.mesh_layers * P.n_mesh_mats * npix * 2;
        for (int i = 0; i < P.n_mesh_mats; i++)
            for (int j = 0; j < P.mesh_layers; j++) {
                size_t idx = ((size_t)j * P.n_mesh_mats + i) * npix * 2 + pix * 2;
                if (fabs((double)add[idx + 1]) < 0.00001) {
                    int mm = P.mesh_mats[i];
                    float v = fmaxf(add[idx], 0.0f);
#pragma unroll
                    for (int m = 0; m < NM; m++) if (m == mm) area[m] = __fadd_rn(area[m], v);
                }
            }
    }
    float* out = P.area + (size_t)view * P.M * npix + pix;
#pragma unroll
    for (int m = 0; m < NM; m++)
        if (!WIDE || m < P.M) out[(size_t)m * npix] = __fdiv_rn(area[m], 10.0f);  // K.cu:582-584
    // S_view bookkeeping
    unsigned long long ns = num_steps > 0 ? (unsigned long long)num_steps * (WIDE ? P.V : NV) : 0ull;
    unsigned mask = __activemask();
    for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(mask, ns, o);
    if ((threadIdx.x & 31) == (__ffs(mask) - 1) && ns) atomicAdd(P.sample_count, ns);
}

template <int NV, int NM>
__global__ void __launch_bounds__(128) march_general_kernel(const __grid_constant__ MarchParams P) {
    const int tiles_x = (P.W + 15) / 16;
    const int tile = blockIdx.x, view = blockIdx.y;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int udx = tx * 16 + (threadIdx.x & 15), vdx = ty * 8 + (threadIdx.x >> 4);
    if (udx >= P.W || vdx >= P.H) return;
    general_ray<NV, NM>(P, view, udx, vdx);
}

// The tiles march_multi_kernel (drr_march_warp.cu) could not take: one warp per 8x4 tile, pulled from the work list.
template <int NV, int NM>
__global__ void __launch_bounds__(128) march_general_list_kernel(const __grid_constant__ MarchParams P) {
    const int lane = threadIdx.x & 31;
    const unsigned n = P.work_count[0];
    const int tiles_x = (P.W + TILE_W - 1) / TILE_W, tiles_y = (P.H + TILE_H - 1) / TILE_H;
    const unsigned tiles_per_view = (unsigned)tiles_x * tiles_y;
    for (;;) {
        unsigned w = 0;
        if (lane == 0) w = atomicAdd(P.work_count + 1, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n) break;
        const unsigned tile = P.worklist[w];
        const unsigned view = tile / tiles_per_view;
        const unsigned tv = tile - view * tiles_per_view;
        const int ty = tv / tiles_x, tx = tv - ty * tiles_x;
        const int udx = tx * TILE_W + (lane & (TILE_W - 1)), vdx = ty * TILE_H + (lane >> 3);
        if (udx < P.W && vdx < P.H) general_ray<NV, NM>(P, (int)view, udx, vdx);
        __syncwarp();
    }
}

// NUM_VOLUMES == 0 (mesh-only scenes): the trace block of projectKernel is compiled out (K.cu:252-555) and
// only the additive mesh densities reach the area densities (K.cu:565-584).
__global__ void march_meshonly_kernel(const __grid_constant__ MarchParams P) {
    const size_t npix = (size_t)P.W * P.H;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * P.n_views) return;
    const size_t view = idx / npix, pix = idx - view * npix;
    float area[DRR_MAX_MATERIALS];
    for (int m = 0; m < P.M; m++) area[m] = 0.0f;
    if (P.additive != nullptr) {
        const float* add = P.additive + view * P.mesh_layers * P.n_mesh_mats * npix * 2;
        for (int i = 0; i < P.n_mesh_mats; i++)
            for (int j = 0; j < P.mesh_layers; j++) {
                size_t k = ((size_t)j * P.n_mesh_mats + i) * npix * 2 + pix * 2;
                if (fabs((double)add[k + 1]) < 0.00001) area[P.mesh_mats[i]] += fmaxf(add[k], 0.0f);
            }
    }
    for (int m = 0; m < P.M; m++) P.area[view * P.M * npix + (size_t)m * npix + pix] = __fdiv_rn(area[m], 10.0f);
}

cudaError_t drr_launch_march_meshonly(const MarchParams& P, cudaStream_t s) {
    size_t total = (size_t)P.W * P.H * P.n_views;
    march_meshonly_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(P);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// launchers (called from drr_capi.cu)
// ---------------------------------------------------------------------------------------------
template <int NM>
static cudaError_t launch_single_nm(const MarchParams& P, int grid, cudaStream_t s) {
    march_single_kernel<NM><<<grid, 256, 0, s>>>(P);
    return cudaGetLastError();
}

cudaError_t drr_launch_march_single(const MarchParams& P, int grid, cudaStream_t s) {
    switch (P.M) {
        case 1: return launch_single_nm<1>(P, grid, s);
        case 2: return launch_single_nm<2>(P, grid, s);
        case 3: return launch_single_nm<3>(P, grid, s);
        case 4: return launch_single_nm<4>(P, grid, s);
        case 5: return launch_single_nm<5>(P, grid, s);
        case 6: return launch_single_nm<6>(P, grid, s);
        case 7: return launch_single_nm<7>(P, grid, s);
        case 8: return launch_single_nm<8>(P, grid, s);
        default: return cudaErrorInvalidValue;
    }
}

template <int NV, int NM>
static cudaError_t launch_general_nvnm(const MarchParams& P, cudaStream_t s) {
    dim3 grid(((P.W + 15) / 16) * ((P.H + 7) / 8), P.n_views);
    march_general_kernel<NV, NM><<<grid, 128, 0, s>>>(P);
    return cudaGetLastError();
}

template <int NV>
static cudaError_t launch_general_nv(const MarchParams& P, cudaStream_t s) {
    switch (P.M) {
        case 1: return launch_general_nvnm<NV, 1>(P, s);
        case 2: return launch_general_nvnm<NV, 2>(P, s);
        case 3: return launch_general_nvnm<NV, 3>(P, s);
        case 4: return launch_general_nvnm<NV, 4>(P, s);
        case 5: return launch_general_nvnm<NV, 5>(P, s);
        case 6: return launch_general_nvnm<NV, 6>(P, s);
        case 7: return launch_general_nvnm<NV, 7>(P, s);
        case 8: return launch_general_nvnm<NV, 8>(P, s);
        default: return cudaErrorInvalidValue;
    }
}

template <int NV, int NM>
static cudaError_t launch_general_list_nvnm(const MarchParams& P, int grid, cudaStream_t s) {
    march_general_list_kernel<NV, NM><<<grid, 128, 0, s>>>(P);
    return cudaGetLastError();
}

template <int NV>
static cudaError_t launch_general_list_nv(const MarchParams& P, int grid, cudaStream_t s) {
    switch (P.M) {
        case 1: return launch_general_list_nvnm<NV, 1>(P, grid, s);
        case 2: return launch_general_list_nvnm<NV, 2>(P, grid, s);
        case 3: return launch_general_list_nvnm<NV, 3>(P, grid, s);
        case 4: return launch_general_list_nvnm<NV, 4>(P, grid, s);
        case 5: return launch_general_list_nvnm<NV, 5>(P, grid, s);
        case 6: return launch_general_list_nvnm<NV, 6>(P, grid, s);
        case 7: return launch_general_list_nvnm<NV, 7>(P, grid, s);
        case 8: return launch_general_list_nvnm<NV, 8>(P, grid, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t drr_launch_march_general_list(const MarchParams& P, int grid, cudaStream_t s) {
    if (P.V > 4 || P.M > 8) return launch_general_list_nvnm<DRR_MAX_VOLUMES, DRR_MAX_MATERIALS>(P, grid, s);
    switch (P.V) {
        case 1: return launch_general_list_nv<1>(P, grid, s);
        case 2: return launch_general_list_nv<2>(P, grid, s);
        case 3: return launch_general_list_nv<3>(P, grid, s);
        case 4: return launch_general_list_nv<4>(P, grid, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t drr_launch_march_general(const MarchParams& P, cudaStream_t s) {
    if (P.V > 4 || P.M > 8) return launch_general_nvnm<DRR_MAX_VOLUMES, DRR_MAX_MATERIALS>(P, s);
    switch (P.V) {
        case 1: return launch_general_nv<1>(P, s);
        case 2: return launch_general_nv<2>(P, s);
        case 3: return launch_general_nv<3>(P, s);
        case 4: return launch_general_nv<4>(P, s);
        default: return cudaErrorInvalidValue;
    }
}

int drr_march_single_occupancy(int M) {
    int nb = 0;
    switch (M) {
        case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, march_single_kernel<1>, 256, 0); break;
        case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, march_single_kernel<2>, 256, 0); break;
        case 3: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, march_single_kernel<3>, 256, 0); break;
        case 4: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, march_single_kernel<4>, 256, 0); break;
        case 5: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, march_single_kernel<5>, 256, 0); break;
        case 6: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, march_single_kernel<6>, 256, 0); break;
        case 7: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, march_single_kernel<7>, 256, 0); break;
        case 8: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, march_single_kernel<8>, 256, 0); break;
        default: break;
    }
    return nb;
}
