// Device-side building blocks shared by the ray-march kernels of libdrr_b200 (sm_100a).
//
// Everything here follows the arithmetic of the reference kernel
// (/root/reference/deepdrr/projector/project_kernel.cu, "K.cu:n" below) and of the B200 texture unit
// it samples through (model reverse-engineered with tools/tex_probe.cu):
//   * explicit __f*_rn intrinsics pin the fused / unfused pattern the reference's SASS has, so the
//     compiler cannot contract differently here;
//   * hw_trilinear_*() reproduce tex3D<float> (linear filter, clamp, unnormalised coordinates) with
//     its 1.8 fixed-point coordinates and hierarchically rounded integer weights.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/drr_b200.h"

#define DRR_MAGIC 12582912.0f /* 1.5 * 2^23: adding it rounds a float in [0, 2^22) to an integer */

struct VolDev {
    const float* dens;    // [nk][nj][ni] density, i fastest (texture order, projector.py:1468-1470)
    const uint8_t* lab;   // [nk][nj][ni] labels (global material index)
    const float4* cellc;  // [(nk+1)][(nj+1)][(ni+1)][2] per-cell filter coefficients (ALU sampler), z-slices interleaved
    const uint2* celll;   // [(nk+1)][(nj+1)][(ni+1)]    per-cell 8 corner labels
    const uint8_t* cellcode;  // same shape: the label if all 8 corners agree and the cell is interior, else 0xFF
    cudaTextureObject_t tex;
    int ni, nj, nk;
    int pad;
};

struct ViewDev {  // per-view inputs, projector.py:802-831
    float w2i[9];
    float src[DRR_MAX_VOLUMES][3];
    float ijk[DRR_MAX_VOLUMES][12];
    float w2i_inv[9];  // inverse of w2i: world vector -> homogeneous pixel (mesh binning only)
};

struct MarchParams {
    VolDev vol[DRR_MAX_VOLUMES];
    int priority[DRR_MAX_VOLUMES];
    int enabled[DRR_MAX_VOLUMES];
    int V, M, W, H, n_views;
    float step, max_ray_length;
    int attenuate_outside, air_index;
    int tex_eighths;  // hybrid sampler: how many of every 8 consecutive steps of a fast segment use the texture unit
    float slack_lo, slack_hi, slack_alpha;  // lock-step kernels: slack of the staged box (voxels) and of the window tests (mm)
    int lane_quads;                         // march_warp_kernel: lanes 4i .. 4i+3 walk a 2 x 2 block of pixels (else a 4 x 1 run)
    int rays_per_lane;                      // march_warp_kernel: 1 or 2 (host only: picks the instantiation)
    // meshes (K.cu:172-177); null when unused
    int mesh_layers, max_hits, n_mesh_mats;
    const float* hit_alphas;
    const int8_t* hit_facing;
    const int8_t* layer_valid;
    const float* additive;
    const int* mesh_mats;
    const ViewDev* views;
    float* area;  // [view][M][H*W]
    unsigned long long* sample_count;
    unsigned int* tile_counter;
    // multi-volume scenes: 8x4-pixel tiles the lock-step kernel hands over to the general kernel
    unsigned int* worklist;    // [n_tiles] tile ids
    unsigned int* work_count;  // [0] entries in the list, [1] consumer cursor
};

// ---------------------------------------------------------------------------------------------
// Ray set-up and slab test (K.cu:220-234, 265-323)
// ---------------------------------------------------------------------------------------------
struct Ray {
    float rx, ry, rz, ray_length;
};

__device__ __forceinline__ Ray make_ray(const float* __restrict__ w, int udx, int vdx) {
    float u = (float)udx + 0.5f, v = (float)vdx + 0.5f;
    // nvcc fuses "u*w0 + v*w1 + w2" as fma(u, w0, v*w1) + w2 in the reference's SASS
    float rx = __fadd_rn(__fmaf_rn(u, w[0], __fmul_rn(v, w[1])), w[2]);
    float ry = __fadd_rn(__fmaf_rn(u, w[3], __fmul_rn(v, w[4])), w[5]);
    float rz = __fadd_rn(__fmaf_rn(u, w[6], __fmul_rn(v, w[7])), w[8]);
    float len = __fsqrt_rn(__fmaf_rn(rz, rz, __fmaf_rn(rx, rx, __fmul_rn(ry, ry))));
    float inv = __frcp_rn(len);  // 1.0f / len, IEEE
    Ray r;
    r.rx = __fmul_rn(rx, inv);
    r.ry = __fmul_rn(ry, inv);
    r.rz = __fmul_rn(rz, inv);
    r.ray_length = len;
    return r;
}

__device__ __forceinline__ void ray_dir_ijk(const Ray& r, const float* __restrict__ A, float& dx, float& dy, float& dz) {
    dx = __fmaf_rn(0.0f, A[3], __fmaf_rn(r.rz, A[2], __fmaf_rn(r.rx, A[0], __fmul_rn(r.ry, A[1]))));
    dy = __fmaf_rn(0.0f, A[7], __fmaf_rn(r.rz, A[6], __fmaf_rn(r.rx, A[4], __fmul_rn(r.ry, A[5]))));
    dz = __fmaf_rn(0.0f, A[11], __fmaf_rn(r.rz, A[10], __fmaf_rn(r.rx, A[8], __fmul_rn(r.ry, A[9]))));
}

// Returns do_trace; lo/hi are the volume's entry / exit alphas (K.cu:283-315).
__device__ __forceinline__ bool slab_test(float dx, float dy, float dz, float sx, float sy, float sz, int ni, int nj, int nk,
                                          float max_ray_length, float& lo, float& hi) {
    lo = 0.0f;
    hi = max_ray_length > 0 ? max_ray_length : INFINITY;
    const float d[3] = {dx, dy, dz}, s[3] = {sx, sy, sz};
    const float mx[3] = {(float)ni - 0.5f, (float)nj - 0.5f, (float)nk - 0.5f};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (0.0f != d[a]) {
            float reci = __frcp_rn(d[a]);
            float a0 = __fmul_rn(__fsub_rn(-0.5f, s[a]), reci);
            float a1 = __fmul_rn(__fsub_rn(mx[a], s[a]), reci);
            lo = fmaxf(lo, fminf(a0, a1));
            hi = fminf(hi, fmaxf(a0, a1));
        } else if (-0.5f > s[a] || s[a] > mx[a]) {
            return false;
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// alpha after n sequential fp32 additions of step -- exactly what n times "alpha += step" (K.cu:552) gives, without
// doing them one by one.  Inside one binade every partial sum lands on the same ulp grid, so each addition moves alpha
// by the same whole number of ulps D (step rounded to that grid; the exact sum stays below the top of the binade, so the
// grid cannot change under it): m additions are one integer multiply-add on the bit pattern.  Additions that leave the
// binade, and binades where step falls exactly between two grid points (round-to-even then depends on the parity of
// alpha), are taken one at a time.  alpha > 0 normal, step > 0.  Checked against the plain loop on 460 000 random
// (alpha, step, n) on the host, ties and powers of two included.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float alpha_jump(float alpha, float step, int n) {
    while (n > 0) {
        const float next = __fadd_rn(alpha, step);
        const unsigned ua = __float_as_uint(alpha), un = __float_as_uint(next);
        if ((ua >> 23) != (un >> 23) || n == 1) { alpha = next; n--; continue; }
        const unsigned D = un - ua;
        if (D == 0u) return alpha;  // step is below half an ulp: alpha no longer moves
        const float r = __fsub_rn(step, __fsub_rn(next, alpha));             // both subtractions are exact
        const float ulp = __uint_as_float(((ua >> 23) - 23u) << 23);
        unsigned m = (0x7FFFFFu - (ua & 0x7FFFFFu)) / D;                      // additions that stay inside the binade
        if (__fmul_rn(fabsf(r), 2.0f) == ulp || m == 0u) { alpha = next; n--; continue; }
        m = min(m, (unsigned)n);
        alpha = __uint_as_float(ua + m * D);
        n -= (int)m;
    }
    return alpha;
}

// ---------------------------------------------------------------------------------------------
// Texture-unit arithmetic, integer form (general / boundary path)
// ---------------------------------------------------------------------------------------------
// Q = clamp(floor((c - 0.5) * 256 + 0.5), 0, (n - 1) * 256) for texture coordinate c.
__device__ __forceinline__ int hw_fix8(float c, int n) {
    float v = __fmaf_rn(__fsub_rn(c, 0.5f), 256.0f, 0.5f);
    int q = (int)floorf(v);
    return max(0, min(q, (n - 1) * 256));
}

// The eight integer weights (sum 256) for fractions a (x), b (y), c (z); w[z][x][y].
__device__ __forceinline__ void hw_weights(int a, int b, int c, int w[2][2][2]) {
#pragma unroll
    for (int zz = 0; zz < 2; zz++) {
        int wz = zz ? c : 256 - c;
        int X1 = (wz * a + 128) >> 8, X0 = wz - X1;
        int Y11 = (X1 * b + 128) >> 8, Y10 = X1 - Y11;
        int Y00 = (X0 * (256 - b) + 128) >> 8, Y01 = X0 - Y00;
        w[zz][0][0] = Y00; w[zz][0][1] = Y01; w[zz][1][0] = Y10; w[zz][1][1] = Y11;
    }
}

// tex3D<float>(vol, cx, cy, cz) emulated from the raw [k][j][i] array.
__device__ __forceinline__ float hw_trilinear_raw(const VolDev& v, float cx, float cy, float cz) {
    int qa = hw_fix8(cx, v.ni), qb = hw_fix8(cy, v.nj), qc = hw_fix8(cz, v.nk);
    int i = qa >> 8, j = qb >> 8, k = qc >> 8;
    int i1 = min(i + 1, v.ni - 1), j1 = min(j + 1, v.nj - 1), k1 = min(k + 1, v.nk - 1);
    int w[2][2][2];
    hw_weights(qa & 255, qb & 255, qc & 255, w);
    const size_t sj = (size_t)v.ni, sk = (size_t)v.ni * v.nj;
    const float* p0 = v.dens + (size_t)k * sk;
    const float* p1 = v.dens + (size_t)k1 * sk;
    // exact sum of products (weights <= 256 and 24-bit texels: every product fits a double exactly)
    double acc = (double)w[0][0][0] * __ldg(p0 + j * sj + i) + (double)w[0][0][1] * __ldg(p0 + j1 * sj + i) +
                 (double)w[0][1][0] * __ldg(p0 + j * sj + i1) + (double)w[0][1][1] * __ldg(p0 + j1 * sj + i1) +
                 (double)w[1][0][0] * __ldg(p1 + j * sj + i) + (double)w[1][0][1] * __ldg(p1 + j1 * sj + i) +
                 (double)w[1][1][0] * __ldg(p1 + j * sj + i1) + (double)w[1][1][1] * __ldg(p1 + j1 * sj + i1);
    return (float)(acc * (1.0 / 256.0));
}

// Point-sampled label fetch tex3D<int>(seg, x, y, z): T[clamp(floor(x))] (K.cu:424).
__device__ __forceinline__ int label_at(const VolDev& v, int i, int j, int k) {
    i = max(0, min(i, v.ni - 1));
    j = max(0, min(j, v.nj - 1));
    k = max(0, min(k, v.nk - 1));
    return __ldg(v.lab + ((size_t)k * v.nj + j) * v.ni + i);
}

// ---------------------------------------------------------------------------------------------
// Texture-unit arithmetic, float form on per-cell records (interior cells, ALU sampler hot path)
// ---------------------------------------------------------------------------------------------
// Cell record of cell base (bi, bj, bk): for z-slice s (k = bk + s) with texels Txy (x = i offset,
// y = j offset), c[s] = (T01, T10 - T01, T00 - T01, T11 - T10) / 256.  With the hardware weights
// X1 = w(x=1), Y00 = w(0,0), Y11 = w(1,1) the slice contribution is
//   wz*T01 + X1*(T10 - T01) + Y00*(T00 - T01) + Y11*(T11 - T10).
//
// fr = fractional cell coordinate in [0, 1) (exact: p - floor(p)).  Returns acc + filtered density.
// Rounding tricks (all exact):
//   RHU(fr*256)      = RN(fr*256 + 2^-15 + MAGIC) - MAGIC          (fr has >= 2^-23 resolution)
//   RHU(P*q/256)     = RN(P * (q/256 + 2^-17) + MAGIC) - MAGIC      (bias P*2^-17 in (0, 2^-8))
__device__ __forceinline__ float hw_trilinear_cell(float xr, float yr, float zr, const float4& c0, const float4& c1, float acc) {
    const float C = DRR_MAGIC;
    float af = __fsub_rn(__fadd_rn(__fmaf_rn(xr, 256.0f, 0x1p-15f), C), C);
    float bf = __fsub_rn(__fadd_rn(__fmaf_rn(yr, 256.0f, 0x1p-15f), C), C);
    float cf = __fsub_rn(__fadd_rn(__fmaf_rn(zr, 256.0f, 0x1p-15f), C), C);
    float wz0 = __fsub_rn(256.0f, cf);
    float wp1 = __fmaf_rn(cf, 0x1p-8f, 0x1p-17f);           // wz1/256 + 2^-17
    float wp0 = __fmaf_rn(cf, -0x1p-8f, 1.0f + 0x1p-17f);   // wz0/256 + 2^-17
    // x fraction rounded up to 256: the hardware is already in the next cell (a = 0) and splits this
    // texel column with the x = 0 rule (y0 rounded half up, y1 the complement), i.e. y1 rounds half
    // DOWN: flip the sign of the tie-breaking bias.
    float bp = __fmaf_rn(bf, 0x1p-8f, af == 256.0f ? -0x1p-17f : 0x1p-17f);  // b/256 +- 2^-17
    float bq = __fmaf_rn(bf, -0x1p-8f, 1.0f + 0x1p-17f);    // (256-b)/256 + 2^-17
    // slice 0
    float X1 = __fsub_rn(__fmaf_rn(wp0, af, C), C);
    float X0 = __fsub_rn(wz0, X1);
    float Y11 = __fsub_rn(__fmaf_rn(X1, bp, C), C);
    float Y00 = __fsub_rn(__fmaf_rn(X0, bq, C), C);
    acc = __fmaf_rn(wz0, c0.x, acc);
    acc = __fmaf_rn(X1, c0.y, acc);
    acc = __fmaf_rn(Y00, c0.z, acc);
    acc = __fmaf_rn(Y11, c0.w, acc);
    // slice 1
    X1 = __fsub_rn(__fmaf_rn(wp1, af, C), C);
    X0 = __fsub_rn(cf, X1);
    Y11 = __fsub_rn(__fmaf_rn(X1, bp, C), C);
    Y00 = __fsub_rn(__fmaf_rn(X0, bq, C), C);
    acc = __fmaf_rn(cf, c1.x, acc);
    acc = __fmaf_rn(X1, c1.y, acc);
    acc = __fmaf_rn(Y00, c1.z, acc);
    acc = __fmaf_rn(Y11, c1.w, acc);
    return acc;
}

// Same arithmetic with the two z-slices packed into f32x2 lanes (FFMA2 / FADD2 / FMUL2, sm_100+): the
// kernel is issue-bound, so halving the instruction count of the slice-symmetric part pays even though
// the FMA pipe does the same work.  Record layout for this form (built when a cell is staged into
// shared memory): A = (c0.x, c1.x, c0.y, c1.y), B = (c0.z, c1.z, c0.w, c1.w).
__device__ __forceinline__ float hw_trilinear_cell2(float xr, float yr, float zr, const float4& A, const float4& B) {
    const float C = DRR_MAGIC;
    const float af = __fsub_rn(__fadd_rn(__fmaf_rn(xr, 256.0f, 0x1p-15f), C), C);
    const float bf = __fsub_rn(__fadd_rn(__fmaf_rn(yr, 256.0f, 0x1p-15f), C), C);
    const float cf = __fsub_rn(__fadd_rn(__fmaf_rn(zr, 256.0f, 0x1p-15f), C), C);
    const float bp = __fmaf_rn(bf, 0x1p-8f, af == 256.0f ? -0x1p-17f : 0x1p-17f);
    const float bq = __fmaf_rn(bf, -0x1p-8f, 1.0f + 0x1p-17f);
    const float2 cc = make_float2(cf, cf), CC = make_float2(C, C), nCC = make_float2(-C, -C);
    const float2 wz = __ffma2_rn(cc, make_float2(-1.0f, 1.0f), make_float2(256.0f, 0.0f));                        // (256 - c, c)
    const float2 wp = __ffma2_rn(cc, make_float2(-0x1p-8f, 0x1p-8f), make_float2(1.0f + 0x1p-17f, 0x1p-17f));      // wz/256 + 2^-17
    const float2 X1 = __fadd2_rn(__ffma2_rn(wp, make_float2(af, af), CC), nCC);
    const float2 X0 = __ffma2_rn(X1, make_float2(-1.0f, -1.0f), wz);
    const float2 Y11 = __fadd2_rn(__ffma2_rn(X1, make_float2(bp, bp), CC), nCC);
    const float2 Y00 = __fadd2_rn(__ffma2_rn(X0, make_float2(bq, bq), CC), nCC);
    float2 r = __fmul2_rn(wz, make_float2(A.x, A.y));
    r = __ffma2_rn(X1, make_float2(A.z, A.w), r);
    r = __ffma2_rn(Y00, make_float2(B.x, B.y), r);
    r = __ffma2_rn(Y11, make_float2(B.z, B.w), r);
    return __fadd_rn(r.x, r.y);
}

// The same filter from the texture unit's own fixed-point coordinates.  q* = 0x4B000000 + Q with Q = RHU(256 * l) the 1.8
// fixed-point coordinate relative to the staged box (one FFMA.RM per axis, see march_core): the low byte is the
// fraction a in [0, 255], the next byte the cell.  A fraction that rounds up to 256 has already moved Q into the next
// cell with a = 0 -- exactly what the unit does -- so the x = 0 rule needs no special case here.
__device__ __forceinline__ float hw_trilinear_cell2q(unsigned qx, unsigned qy, unsigned qz, const float4& A, const float4& B) {
    const float C = DRR_MAGIC;
    const float af = (float)(qx & 0xFFu), bf = (float)(qy & 0xFFu), cf = (float)(qz & 0xFFu);  // I2FP: off the FMA pipes, which bind this loop
    const float bp = __fmaf_rn(bf, 0x1p-8f, 0x1p-17f);
    const float bq = __fmaf_rn(bf, -0x1p-8f, 1.0f + 0x1p-17f);
    const float2 cc = make_float2(cf, cf), CC = make_float2(C, C), nCC = make_float2(-C, -C);
    const float2 wz = make_float2(__fsub_rn(256.0f, cf), cf);                                                     // (256 - c, c)
    const float2 wp = __ffma2_rn(cc, make_float2(-0x1p-8f, 0x1p-8f), make_float2(1.0f + 0x1p-17f, 0x1p-17f));      // wz/256 + 2^-17
    const float2 X1 = __fadd2_rn(__ffma2_rn(wp, make_float2(af, af), CC), nCC);
    const float2 X0 = __ffma2_rn(X1, make_float2(-1.0f, -1.0f), wz);
    const float2 Y11 = __fadd2_rn(__ffma2_rn(X1, make_float2(bp, bp), CC), nCC);
    const float2 Y00 = __fadd2_rn(__ffma2_rn(X0, make_float2(bq, bq), CC), nCC);
    float2 r = __fmul2_rn(wz, make_float2(A.x, A.y));
    r = __ffma2_rn(X1, make_float2(A.z, A.w), r);
    r = __ffma2_rn(Y00, make_float2(B.x, B.y), r);
    r = __ffma2_rn(Y11, make_float2(B.z, B.w), r);
    return __fadd_rn(r.x, r.y);
}

// Trilinear one-hot material weights of the reference (K.cu:434-455), full fp32 weights.
// lab8: byte (dx + 2*dy + 4*dz) = label of corner (dx, dy, dz).  seg[] must be zeroed by the caller.
template <int NM>
__device__ __forceinline__ void seg_weights(float fx, float fy, float fz, uint2 lab8, float* seg) {
    float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy), gz = __fsub_rn(1.0f, fz);
    // accumulation order of the reference: dz outer, dy, dx inner
#pragma unroll
    for (int c = 0; c < 2; c++) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
#pragma unroll
            for (int a = 0; a < 2; a++) {
                float w = __fmul_rn(__fmul_rn(a ? fx : gx, b ? fy : gy), c ? fz : gz);
                unsigned word = c ? lab8.y : lab8.x;
                int l = (int)__byte_perm(word, 0u, 0x4440u + (a + 2 * b));
                // one predicated add per material (adding 0.0f to the others, as a select would, changes nothing); written
                // in PTX because the compiler turns the C++ form into select + add or into branches
#pragma unroll
                for (int m = 0; m < NM; m++)
                    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\t@p add.rn.f32 %0, %0, %1;\n\t}" : "+f"(seg[m]) : "f"(w), "r"(l), "r"(m));
            }
        }
    }
}
