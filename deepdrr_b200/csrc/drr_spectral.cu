// Spectral Beer-Lambert + detector post-processing kernels of libdrr_b200 (sm_100a).
//
// Replace the tail of the reference's `projectKernel` (project_kernel.cu:637-646, "K.cu:n") and the host
// post-processing of `Projector.project` (projector.py:691-702):
//   spectral_kernel   intensity = sum_E E * pdf(E) * exp(-sum_m mu_m(E) * A_m), photon_prob likewise
//                     (tables staged in shared memory, sums kept in registers instead of the
//                     reference's global read-modify-write per bin)
//   noise kernels     analytic_generators.add_noise (analytic_generators.py:10-18): Poisson shot noise
//                     (Philox), 3x3 blur, clip at 0
//   clip              np.clip(images, None, intensity_upper_bound)            (projector.py:697-698)
//   minmax + neglog   utils.neglog (utils/image_utils.py:18-59): per-image min / max by warp shuffles
//                     + one atomic per block, then -log(I + min + eps) scaled to [0, 1]
//   solid angle       calculate_solid_angle + _calculate_collected_energy_per_pixel
//                     (K.cu:14-133, projector.py:833-853)
// None of this is a dense contraction: it runs on the SIMT pipes (FMA + MUFU.EX2).
#include <curand_kernel.h>
#include <math_constants.h>

#include "drr_device.cuh"

// order-preserving float <-> uint key (works for negative values as well)
__device__ __forceinline__ unsigned f2key(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    unsigned b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(b);
}

// exp2 on the SFU (MUFU.EX2): 2 ulp of the result; arguments below -126 give 0, like the denormals `expf` would return
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Per energy bin one record in shared memory: mu_m(E) * (-log2 e) for every material, pdf(E), E * pdf(E), padded to a multiple
// of four floats (two 128-bit broadcast loads for up to six materials).  The bin is then  ex = 2^(sum_m A_m mu'_m)  (one MUFU),
// photon_prob += ex * pdf, intensity += ex * (E pdf): 3 + NM instructions instead of the ~25 of `expf` and separate tables.
// The reference's form (K.cu:637-646: expf, p = exp * pdf, intensity += E * p) differs by the roundings of the pre-scaled table
// and of ex2.approx: <= 3e-6 relative at an exponent of 30, against the 1e-4 north_star allows on intensity
// (tests/test_gpu_parity.py measures it against the reference-kernel goldens).
__host__ __device__ constexpr int spectral_stride(int M) { return (M + 2 + 3) & ~3; }

template <int NM>
__global__ void __launch_bounds__(256) spectral_kernel(const float* __restrict__ area, int n_bins, int M_rt,
                                                       const float* __restrict__ energies, const float* __restrict__ pdf,
                                                       const float* __restrict__ mu, size_t npix, int n_views,
                                                       float* __restrict__ intensity, float* __restrict__ pprob) {
    extern __shared__ __align__(16) float sm[];
    const int M = NM > 0 ? NM : M_rt;
    const int S = NM > 0 ? spectral_stride(NM) : spectral_stride(M_rt);
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) {
        float* r = sm + (size_t)i * S;
        for (int m = 0; m < M; m++) r[m] = __fmul_rn(mu[i * M + m], -1.4426950408889634f);
        for (int m = M; m < S - 2; m++) r[m] = 0.0f;
        r[S - 2] = pdf[i];
        r[S - 1] = __fmul_rn(energies[i], pdf[i]);
    }
    __syncthreads();
    const size_t total = npix * (size_t)n_views;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t view = idx / npix, pix = idx - view * npix;
        const float* a = area + view * (size_t)M * npix + pix;
        float A[NM > 0 ? NM : DRR_MAX_MATERIALS];
#pragma unroll
        for (int m = 0; m < (NM > 0 ? NM : DRR_MAX_MATERIALS); m++) A[m] = (m < M) ? a[(size_t)m * npix] : 0.0f;
        float inten = 0.0f, pp = 0.0f;
#pragma unroll 4
        for (int b = 0; b < n_bins; b++) {
            const float* r = sm + (size_t)b * S;
            float e = 0.0f;
#pragma unroll
            for (int m = 0; m < (NM > 0 ? NM : DRR_MAX_MATERIALS); m++)
                if (m < M) e = __fmaf_rn(A[m], r[m], e);
            const float ex = ex2_approx(e);
            pp = __fmaf_rn(ex, r[S - 2], pp);
            inten = __fmaf_rn(ex, r[S - 1], inten);
        }
        intensity[idx] = inten;
        if (pprob) pprob[idx] = pp;
    }
}

cudaError_t drr_launch_spectral(const float* area, int n_bins, int M, const float* energies, const float* pdf, const float* mu,
                                size_t npix, int n_views, float* intensity, float* pprob, int n_sm, cudaStream_t s) {
    size_t total = npix * (size_t)n_views;
    int grid = (int)((total + 255) / 256);
    if (grid > n_sm * 8) grid = n_sm * 8;
    size_t smem = sizeof(float) * (size_t)n_bins * spectral_stride(M);
#define SPEC(N) spectral_kernel<N><<<grid, 256, smem, s>>>(area, n_bins, M, energies, pdf, mu, npix, n_views, intensity, pprob)
    switch (M) {
        case 1: SPEC(1); break;
        case 2: SPEC(2); break;
        case 3: SPEC(3); break;
        case 4: SPEC(4); break;
        case 5: SPEC(5); break;
        case 6: SPEC(6); break;
        default: SPEC(0); break;
    }
#undef SPEC
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// noise (analytic_generators.py:10-18)
// ---------------------------------------------------------------------------------------------
__global__ void shot_noise_kernel(const float* __restrict__ intensity, const float* __restrict__ pprob, float photon_count,
                                  size_t total, unsigned long long seed, float* __restrict__ shot) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, idx, 0, &st);
    double lam = (double)pprob[idx] * (double)photon_count;  // np.random.poisson works in float64
    double k = (double)curand_poisson(&st, lam);
    shot[idx] = (float)((k - lam) * (double)intensity[idx] / lam);
}

__global__ void blur_add_clip_kernel(const float* __restrict__ shot, int W, int H, int n_views, float* __restrict__ intensity) {
    // scipy.signal.convolve2d(shot, kernel_shot_noise, mode="same"): true convolution, zero padding
    const float k[3][3] = {{0.03f, 0.06f, 0.02f}, {0.11f, 0.98f, 0.11f}, {0.02f, 0.06f, 0.03f}};
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t npix = (size_t)W * H;
    if (idx >= npix * n_views) return;
    size_t view = idx / npix, pix = idx - view * npix;
    int y = (int)(pix / W), x = (int)(pix - (size_t)y * W);
    const float* s = shot + view * npix;
    double acc = 0.0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            int yy = y - dy, xx = x - dx;  // convolution: out[y,x] = sum k[dy+1][dx+1] * in[y-dy, x-dx]
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc += (double)k[dy + 1][dx + 1] * (double)s[(size_t)yy * W + xx];
        }
    double v = (double)intensity[idx] + acc;
    intensity[idx] = (float)fmin(fmax(v, 0.0), 10e30);
}

__global__ void clip_upper_kernel(float* __restrict__ img, size_t total, float upper) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < total) img[idx] = fminf(img[idx], upper);
}

// ---------------------------------------------------------------------------------------------
// neglog (utils/image_utils.py:18-59)
// ---------------------------------------------------------------------------------------------
// minmax[2*view] = key(min), minmax[2*view+1] = key(max); must be initialised to 0xFFFFFFFF / 0.
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ img, size_t npix, unsigned* __restrict__ minmax) {
    const int view = blockIdx.y;
    const float* p = img + (size_t)view * npix;
    unsigned lo = 0xFFFFFFFFu, hi = 0u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        unsigned k = f2key(p[i]);
        lo = min(lo, k);
        hi = max(hi, k);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ unsigned s_lo[8], s_hi[8];
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
        atomicMin(minmax + 2 * view, lo);
        atomicMax(minmax + 2 * view + 1, hi);
    }
}

__global__ void __launch_bounds__(256) neglog_kernel(float* __restrict__ img, size_t npix, int n_views,
                                                     const unsigned* __restrict__ minmax, float epsilon, int* __restrict__ const_flag) {
    // "if np.any(image_max == image_min): image[:] = 0" zeroes the whole batch (image_utils.py:42-49)
    __shared__ int s_const;
    if (threadIdx.x == 0) s_const = 0;
    __syncthreads();
    for (int v = threadIdx.x; v < n_views; v += blockDim.x) {
        float mn = key2f(minmax[2 * v]), mx = key2f(minmax[2 * v + 1]);
        float shift = __fadd_rn(mn, epsilon);
        float a = -logf(__fadd_rn(mx, shift)), b = -logf(__fadd_rn(mn, shift));
        if (a == b) atomicOr(&s_const, 1);
    }
    __syncthreads();
    const bool all_zero = s_const != 0;
    // a batch projected in pieces (drr_project's copy / compute pipeline) learns here that one piece held a constant image
    if (all_zero && const_flag != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *const_flag = 1;
    const int view = blockIdx.y;
    const float mn = key2f(minmax[2 * view]), mx = key2f(minmax[2 * view + 1]);
    const float shift = __fadd_rn(mn, epsilon);
    const float lo = -logf(__fadd_rn(mx, shift)), hi = -logf(__fadd_rn(mn, shift));
    const float d = __fsub_rn(hi, lo);
    float* p = img + (size_t)view * npix;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        float v = -logf(__fadd_rn(p[i], shift));
        p[i] = all_zero ? 0.0f : __fdiv_rn(__fsub_rn(v, lo), d);
    }
}

// ---------------------------------------------------------------------------------------------
// collected energy (K.cu:14-133, projector.py:833-853)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float solid_angle_px(const float* __restrict__ w, int udx, int vdx) {
    // Explicit roundings: the contraction pattern nvcc gives the reference's calculate_solid_angle (read from the SASS of the
    // unmodified kernel).  The result is a difference of nearly equal products, so a different pattern moves it by 1e-4 relative.
    float cx[4], cy[4], cz[4], cm[4];
    const float cu_off[4] = {0.f, 1.f, 1.f, 0.f}, cv_off[4] = {0.f, 0.f, 1.f, 1.f};
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float cu = __fadd_rn((float)udx, cu_off[c]), cv = __fadd_rn((float)vdx, cv_off[c]);
        cx[c] = __fadd_rn(__fmaf_rn(w[0], cu, __fmul_rn(w[1], cv)), w[2]);
        cy[c] = __fadd_rn(__fmaf_rn(w[3], cu, __fmul_rn(w[4], cv)), w[5]);
        cz[c] = __fadd_rn(__fmaf_rn(w[6], cu, __fmul_rn(w[7], cv)), w[8]);
        cm[c] = __fsqrt_rn(__fmaf_rn(cz[c], cz[c], __fmaf_rn(cx[c], cx[c], __fmul_rn(cy[c], cy[c]))));
    }
    auto dot = [&](int a, int b) { return __fmaf_rn(cz[a], cz[b], __fmaf_rn(cx[a], cx[b], __fmul_rn(cy[a], cy[b]))); };
    const float kx = __fmaf_rn(cy[0], cz[2], -__fmul_rn(cz[0], cy[2]));
    const float ky = __fmaf_rn(cz[0], cx[2], -__fmul_rn(cx[0], cz[2]));
    const float kz = __fmaf_rn(cx[0], cy[2], -__fmul_rn(cy[0], cx[2]));
    const float d01 = dot(0, 1), d02 = dot(0, 2), d03 = dot(0, 3), d12 = dot(1, 2), d23 = dot(2, 3);
    const float n012 = fabsf(__fmaf_rn(cz[1], kz, __fmaf_rn(cx[1], kx, __fmul_rn(cy[1], ky))));
    const float n023 = fabsf(__fmaf_rn(cz[3], kz, __fmaf_rn(cx[3], kx, __fmul_rn(cy[3], ky))));
    const float e012 = __fmaf_rn(d12, cm[0], __fmaf_rn(d02, cm[1], __fmaf_rn(__fmul_rn(cm[1], cm[0]), cm[2], __fmul_rn(d01, cm[2]))));
    const float e023 = __fmaf_rn(d23, cm[0], __fmaf_rn(d03, cm[2], __fmaf_rn(__fmul_rn(cm[2], cm[0]), cm[3], __fmul_rn(d02, cm[3]))));
    float s1 = atan2f(n012, e012);
    s1 = __fadd_rn(s1, s1);
    if (s1 < 0.0f) s1 = __fadd_rn(s1, CUDART_PI_F);
    float s2 = atan2f(n023, e023);
    s2 = __fadd_rn(s2, s2);
    if (s2 < 0.0f) s2 = __fadd_rn(s2, CUDART_PI_F);
    return __fadd_rn(s1, s2);
}

// pass 1: solid angle per pixel + per-view sum (double); pass 2: scale the intensity
__global__ void __launch_bounds__(256) solid_angle_kernel(const ViewDev* __restrict__ views, int W, int H, float* __restrict__ solid,
                                                          double* __restrict__ view_sum) {
    const int view = blockIdx.y;
    const size_t npix = (size_t)W * H;
    double part = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        int vdx = (int)(i / W), udx = (int)(i - (size_t)vdx * W);
        float sa = solid_angle_px(views[view].w2i, udx, vdx);
        solid[(size_t)view * npix + i] = sa;
        part += (double)sa;
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(view_sum + view, part);
}

__global__ void collected_energy_kernel(float* __restrict__ intensity, const float* __restrict__ solid, const double* __restrict__ view_sum,
                                        size_t npix, int n_views, float photon_count, float pixel_area) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * n_views) return;
    size_t view = idx / npix;
    // deposited = I * omega * photon_count / mean(omega) / (px * py)   (projector.py:846-852, float64 mean)
    double mean = view_sum[view] / (double)npix;
    double dep = (double)__fmul_rn(intensity[idx], solid[idx]) * (double)photon_count / mean;
    intensity[idx] = (float)(dep / (double)pixel_area);
}

// ---------------------------------------------------------------------------------------------
// launch helpers for drr_capi.cu
// ---------------------------------------------------------------------------------------------
cudaError_t drr_launch_noise(float* intensity, const float* pprob, float* scratch, int W, int H, int n_views, float photon_count,
                             unsigned long long seed, cudaStream_t s) {
    size_t total = (size_t)W * H * n_views;
    int grid = (int)((total + 255) / 256);
    shot_noise_kernel<<<grid, 256, 0, s>>>(intensity, pprob, photon_count, total, seed, scratch);
    blur_add_clip_kernel<<<grid, 256, 0, s>>>(scratch, W, H, n_views, intensity);
    return cudaGetLastError();
}

cudaError_t drr_launch_clip(float* img, size_t total, float upper, cudaStream_t s) {
    clip_upper_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(img, total, upper);
    return cudaGetLastError();
}

cudaError_t drr_launch_neglog(float* img, size_t npix, int n_views, unsigned* minmax, float epsilon, cudaStream_t s, int* const_flag) {
    // minmax initial values: (0xFFFFFFFF, 0) per view
    cudaError_t e = cudaMemsetAsync(minmax, 0, sizeof(unsigned) * 2 * n_views, s);
    if (e != cudaSuccess) return e;
    // set the "min" slots to all-ones with a strided 2-D memset
    e = cudaMemset2DAsync(minmax, 2 * sizeof(unsigned), 0xFF, sizeof(unsigned), n_views, s);
    if (e != cudaSuccess) return e;
    int gx = (int)((npix + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    dim3 grid(gx, n_views);
    minmax_kernel<<<grid, 256, 0, s>>>(img, npix, minmax);
    neglog_kernel<<<grid, 256, 0, s>>>(img, npix, n_views, minmax, epsilon, const_flag);
    return cudaGetLastError();
}

cudaError_t drr_launch_collected(float* intensity, float* solid, double* view_sum, const ViewDev* views, int W, int H, int n_views,
                                 float photon_count, float pixel_area, cudaStream_t s) {
    size_t npix = (size_t)W * H;
    cudaError_t e = cudaMemsetAsync(view_sum, 0, sizeof(double) * n_views, s);
    if (e != cudaSuccess) return e;
    int gx = (int)((npix + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    solid_angle_kernel<<<dim3(gx, n_views), 256, 0, s>>>(views, W, H, solid, view_sum);
    size_t total = npix * n_views;
    collected_energy_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(intensity, solid, view_sum, npix, n_views, photon_count, pixel_area);
    return cudaGetLastError();
}
