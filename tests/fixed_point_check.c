/* Host-side check of the integer tricks of the FMA-pipe sampler (deepdrr_b200/csrc/drr_march_warp.cu: march_core,
 * deepdrr_b200/csrc/drr_device.cuh: hw_trilinear_cell2q), restated in plain C with the CPU's IEEE fmaf:
 *
 *  1. the texture unit's 1.8 fixed-point coordinate relative to the staged box from ONE round-down FMA per axis,
 *         bits(fma_rd(x, 256, 2^23 + 0.5 - 256 * b1)) - 0x4B000000 == floor(256 * (x - b1) + 0.5),
 *     and the floor form used for the label cell (constant without the 0.5);
 *  2. the staging index from the three cell bytes: two byte permutes and one dp4a == cx + nx * (cy + ny * cz);
 *  3. the multiply-shift decomposition of the staging loop: (e * m) >> 16 == e / n for m = trunc(65536 / n) + 2;
 *  4. the float form of the trilinear filter evaluated from those coordinates against the texture unit's integer
 *     model (hw_weights in drr_device.cuh, tex_linear in oracle/drr_oracle.c): Q = RHU(256 * (c - 0.5)), weights split
 *     z -> x -> y with half-up rounding, result = sum(w * T) / 256 rounded once.
 *
 * Built and run by tests/test_fixed_point.py; exit code 0 = the exact checks hold (coordinates, indices, divisions, all
 * eight weights for every fraction triple) and the filter value -- an fp32 FMA chain over those exact weights -- stays within
 * four ulps of the largest texel of the once-rounded exact sum (eight roundings of at most half an ulp each). */
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static double urand(void) { return rand() / (RAND_MAX + 1.0); }

/* fma.rm.f32 */
static float fma_rd(float a, float b, float c) {
    volatile float va = a, vb = b, vc = c;
    fesetround(FE_DOWNWARD);
    volatile float r = fmaf(va, vb, vc);
    fesetround(FE_TONEAREST);
    return r;
}
static float fma_rn(float a, float b, float c) { volatile float r = fmaf(a, b, c); return r; }
static float add_rn(float a, float b) { volatile float r = a + b; return r; }
static float mul_rn(float a, float b) { volatile float r = a * b; return r; }

/* __byte_perm(a, b, sel): result byte i = byte (sel >> 4i) & 7 of the pair {b, a} (selector msb = sign replicate, unused here) */
static uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
    uint64_t pair = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((pair >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static uint32_t dp4a_u8(uint32_t a, uint32_t b, uint32_t c) {
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF);
    return c;
}

/* the unit's integer weights (drr_device.cuh: hw_weights), w[z][x][y] */
static void hw_weights(int a, int b, int c, int w[2][2][2]) {
    for (int zz = 0; zz < 2; zz++) {
        int wz = zz ? c : 256 - c;
        int X1 = (wz * a + 128) >> 8, X0 = wz - X1;
        int Y11 = (X1 * b + 128) >> 8, Y10 = X1 - Y11;
        int Y00 = (X0 * (256 - b) + 128) >> 8, Y01 = X0 - Y00;
        w[zz][0][0] = Y00; w[zz][0][1] = Y01; w[zz][1][0] = Y10; w[zz][1][1] = Y11;
    }
}

/* hw_trilinear_cell2q, one z-slice at a time (the device packs the two slices into f32x2 lanes: same operations).
 * A = (c0.x, c1.x, c0.y, c1.y), B = (c0.z, c1.z, c0.w, c1.w) with c[s] = (T01, T10 - T01, T00 - T01, T11 - T10) / 256. */
static float cell2q(uint32_t qx, uint32_t qy, uint32_t qz, const float A[4], const float B[4]) {
    const float C = 12582912.0f;
    float af = (float)(qx & 0xFF), bf = (float)(qy & 0xFF), cf = (float)(qz & 0xFF);
    float bp = fma_rn(bf, 0x1p-8f, 0x1p-17f), bq = fma_rn(bf, -0x1p-8f, 1.0f + 0x1p-17f);
    float wz[2] = {add_rn(256.0f, -cf), cf};
    float wp[2] = {fma_rn(cf, -0x1p-8f, 1.0f + 0x1p-17f), fma_rn(cf, 0x1p-8f, 0x1p-17f)};
    float r[2];
    for (int s = 0; s < 2; s++) {
        float X1 = add_rn(fma_rn(wp[s], af, C), -C);
        float X0 = fma_rn(X1, -1.0f, wz[s]);
        float Y11 = add_rn(fma_rn(X1, bp, C), -C);
        float Y00 = add_rn(fma_rn(X0, bq, C), -C);
        r[s] = mul_rn(wz[s], A[s]);
        r[s] = fma_rn(X1, A[2 + s], r[s]);
        r[s] = fma_rn(Y00, B[s], r[s]);
        r[s] = fma_rn(Y11, B[2 + s], r[s]);
    }
    return add_rn(r[0], r[1]);
}

int main(void) {
    srand(7);
    long bad_q = 0, bad_floor = 0, bad_idx = 0, bad_div = 0, n_q = 0;

    /* 1. fixed-point coordinate: random box origins and in-box positions, plus exact ties and values next to them */
    for (long it = 0; it < 3000000; it++) {
        int b1 = 2 + rand() % 600;                         /* box origin + 1; >= 2 on interior cells */
        float l;
        int kind = it % 4;
        if (kind == 0) l = (float)(urand() * 12.0);
        else if (kind == 1) l = (float)((rand() % 3072) / 256.0 + 0.5 / 256.0);                  /* ties: 256 l + 0.5 is an integer */
        else if (kind == 2) l = nextafterf((float)((rand() % 3072) / 256.0 + 0.5 / 256.0), (it & 4) ? 100.0f : -100.0f);
        else l = (float)(rand() % 12) + (float)(1.0 - urand() * (1.0 / 400.0));                  /* fractions that round up to 256 */
        volatile float xv = (float)b1 + l;                 /* the sample coordinate as an fp32 number */
        float x = xv;
        double le = (double)x - (double)b1;                /* exact */
        if (le < 0.0 || le >= 250.0) continue;
        float kf = fma_rn(-256.0f, (float)b1, 8388608.0f), kq = add_rn(kf, 0.5f);
        if ((double)kq != 8388608.5 - 256.0 * b1) { bad_q++; continue; }                          /* the constant must be exact */
        uint32_t q = f2u(fma_rd(x, 256.0f, kq)) - 0x4B000000u;
        uint32_t want = (uint32_t)floor(256.0 * le + 0.5);
        if (q != want) { if (bad_q < 5) printf("Q mismatch: x=%.9g b1=%d got %u want %u\n", x, b1, q, want); bad_q++; }
        uint32_t qf = f2u(fma_rd(x, 256.0f, kf)) - 0x4B000000u;
        if ((qf >> 8) != (uint32_t)floor(le) || qf != (uint32_t)floor(256.0 * le)) bad_floor++;
        n_q++;
    }

    /* 2. staging index from the cell bytes, for every box shape march_core's box_fits() lets through: up to 96 cells with the
     * coefficient records, up to 33 * 96 code-only cells for the TEX sampler provided each side and nx * ny fit a byte */
    long n_idx = 0;
    for (long it = 0; it < 2000000; it++) {
        int big = it & 1;
        int nx = 1 + rand() % (big ? 300 : 12), ny = 1 + rand() % (big ? 300 : 12), nz = 1 + rand() % (big ? 40 : 12);
        int cap = big ? 33 * 96 : 96;
        int mx = nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz);
        int fits = big ? (nx * ny * nz <= cap && mx <= 255 && (nx * ny <= 255 || nz == 1)) : (nx * ny * nz <= cap);
        if (!fits) continue;
        int cx = rand() % nx, cy = rand() % ny, cz = rand() % nz;
        uint32_t qx = 0x4B000000u | (cx << 8) | (rand() & 0xFF), qy = 0x4B000000u | (cy << 8) | (rand() & 0xFF), qz = 0x4B000000u | (cz << 8) | (rand() & 0xFF);
        uint32_t cell_w = 1u | ((uint32_t)(nx < 255 ? nx : 255) << 8) | ((uint32_t)(nx * ny < 255 ? nx * ny : 255) << 16);
        uint32_t idx = dp4a_u8(byte_perm(byte_perm(qx, qy, 0x0051), qz, 0x0510), cell_w, 0);
        if (idx != (uint32_t)(cx + nx * (cy + ny * cz))) bad_idx++;
        n_idx++;
    }

    /* 3. multiply-shift division of the staging loop (valid while e * n < 21845) */
    for (int n = 1; n < 400; n++) {
        volatile float inv = 65536.0f / (float)n;
        uint32_t m = (uint32_t)inv + 2u;
        for (uint32_t e = 0; e * (uint32_t)n < 21845u; e++)
            if (((e * m) >> 16) != e / (uint32_t)n) bad_div++;
    }

    /* 4a. the eight weights: every (a, b, c), both z-slices -- the magic-number roundings must reproduce the integer splits */
    long bad_w = 0;
    {
        const float C = 12582912.0f;
        for (int c = 0; c < 256; c++) for (int b = 0; b < 256; b++) for (int a = 0; a < 256; a++) {
            int w[2][2][2];
            hw_weights(a, b, c, w);
            float af = (float)a, bf = (float)b, cf = (float)c;
            float bp = fma_rn(bf, 0x1p-8f, 0x1p-17f), bq = fma_rn(bf, -0x1p-8f, 1.0f + 0x1p-17f);
            float wz[2] = {256.0f - cf, cf};
            float wp[2] = {fma_rn(cf, -0x1p-8f, 1.0f + 0x1p-17f), fma_rn(cf, 0x1p-8f, 0x1p-17f)};
            for (int s = 0; s < 2; s++) {
                float X1 = add_rn(fma_rn(wp[s], af, C), -C), X0 = fma_rn(X1, -1.0f, wz[s]);
                float Y11 = add_rn(fma_rn(X1, bp, C), -C), Y00 = add_rn(fma_rn(X0, bq, C), -C);
                if (X1 != (float)(w[s][1][0] + w[s][1][1]) || X0 != (float)(w[s][0][0] + w[s][0][1]) || Y11 != (float)w[s][1][1] || Y00 != (float)w[s][0][0]) bad_w++;
            }
        }
    }

    /* 4b. the filter value: realistic cells (neighbouring texels within a few per cent, air-like and tissue-like levels, clipped
     * zeros): the float form is an FMA chain over exact weights, so it can differ from the once-rounded exact sum by a few ulps */
    long n_f = 0, n_eq = 0, n_far = 0;
    for (long it = 0; it < 1000000; it++) {
        float T[2][2][2]; /* [z][x][y] */
        float tmax = 0.0f;
        double level = (it % 3 == 0) ? 0.02 * urand() : 0.9 + 1.0 * urand();
        for (int z = 0; z < 2; z++) for (int xx = 0; xx < 2; xx++) for (int y = 0; y < 2; y++) {
            T[z][xx][y] = (float)(level * (1.0 + 0.04 * (urand() - 0.5)));
            if (it % 7 == 0 && (rand() & 3) == 0) T[z][xx][y] = 0.0f;   /* clipped air */
            tmax = fmaxf(tmax, T[z][xx][y]);
        }
        float A[4], B[4];
        const float s = 1.0f / 256.0f;
        for (int z = 0; z < 2; z++) {  /* build_cells_kernel (drr_capi.cu) */
            float t01 = T[z][0][1], t10 = T[z][1][0], t00 = T[z][0][0], t11 = T[z][1][1];
            A[z] = mul_rn(t01, s); A[2 + z] = mul_rn(add_rn(t10, -t01), s);
            B[z] = mul_rn(add_rn(t00, -t01), s); B[2 + z] = mul_rn(add_rn(t11, -t10), s);
        }
        int a = rand() & 255, b = rand() & 255, c = rand() & 255;
        if (it % 11 == 0) a = 0;
        if (it % 13 == 0) b = 0;
        uint32_t qx = 0x4B000000u | (3 << 8) | a, qy = 0x4B000000u | (1 << 8) | b, qz = 0x4B000000u | (2 << 8) | c;
        float got = cell2q(qx, qy, qz, A, B);
        int w[2][2][2];
        hw_weights(a, b, c, w);
        double acc = 0.0;  /* exact: weights <= 256, 24-bit texels */
        for (int z = 0; z < 2; z++) for (int xx = 0; xx < 2; xx++) for (int y = 0; y < 2; y++) acc += (double)w[z][xx][y] * (double)T[z][xx][y];
        float want = (float)(acc / 256.0);
        n_f++;
        if (got == want) n_eq++;
        else {
            float ulp = u2f(f2u(fmaxf(tmax, 1e-30f)) & 0x7F800000u) * 0x1p-23f;   /* ulp of the largest texel */
            if (fabsf(got - want) > 4.0f * ulp) { if (n_far < 5) printf("filter: got %.9g want %.9g (a,b,c)=(%d,%d,%d)\n", got, want, a, b, c); n_far++; }
        }
    }
    printf("fixed-point coordinate: %ld cases, bad %ld, floor form bad %ld; index: %ld boxes, bad %ld; division bad %ld\n", n_q, bad_q, bad_floor, n_idx, bad_idx, bad_div);
    printf("weights: all 256^3 x 2 slices, bad %ld\n", bad_w);
    printf("filter: %ld samples, bit-equal %.4f, beyond four ulps of the largest texel %ld\n", n_f, (double)n_eq / n_f, n_far);
    int ok = bad_q == 0 && bad_floor == 0 && bad_idx == 0 && bad_div == 0 && bad_w == 0 && n_far == 0 && n_eq > 0.5 * n_f;
    printf(ok ? "all ok\n" : "FAILED\n");
    return ok ? 0 : 1;
}
