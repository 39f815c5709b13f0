// Probe of the B200 texture unit's linear-filter arithmetic + a few SIMT throughput microbenchmarks.
// Development tool (not product, not oracle).  Writes raw little-endian arrays under gpurun_out/.
//
//  texprobe_ramp.bin : for axis a in {x,y,z}: 65536 samples of a unit ramp T[i]=i at coordinate
//                      1.5 + k/65536 on that axis (others fixed at 1.5) -> reveals the fixed-point
//                      weight quantisation (CUDA programming guide: 9-bit weights, 8 fractional bits).
//  texprobe_rand.bin : N random 3-D coordinates in a 32^3 random volume: (x,y,z,result) float4, plus
//                      texprobe_vol.bin (the volume, [z][y][x]) to test lerp-arithmetic candidates offline.
//  stdout            : FFMA vs FFMA2 (fma.rn.f32x2) issue throughput, FADD/FSETP mixes.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__global__ void ramp_kernel(cudaTextureObject_t tex, float* out, int axis, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float c = 1.5f + (float)k / (float)n;
    float x = 1.5f, y = 1.5f, z = 1.5f;
    if (axis == 0) x = c;
    if (axis == 1) y = c;
    if (axis == 2) z = c;
    out[k] = tex3D<float>(tex, x, y, z);
}

__global__ void rand_kernel(cudaTextureObject_t tex, const float4* coords, float4* out, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float4 c = coords[k];
    float r = tex3D<float>(tex, c.x, c.y, c.z);
    out[k] = make_float4(c.x, c.y, c.z, r);
}

// ---- throughput microbenchmarks ---------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) fma_bench(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    if (MODE == 0) {  // 8 independent scalar FFMA chains
        for (int i = 0; i < iters; i++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    } else if (MODE == 1) {  // 4 independent packed chains = the same 8 FMAs
        unsigned long long p0, p1, p2, p3, pa, pb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(x0), "f"(x1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(x2), "f"(x3));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(x4), "f"(x5));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(x6), "f"(x7));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
        for (int i = 0; i < iters; i++) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
        }
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(p0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(p1));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x4), "=f"(x5) : "l"(p2));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x6), "=f"(x7) : "l"(p3));
    } else if (MODE == 2) {  // 8 independent FADD chains
        for (int i = 0; i < iters; i++) {
            x0 += a; x1 += a; x2 += a; x3 += a; x4 += a; x5 += a; x6 += a; x7 += a;
        }
    } else if (MODE == 3) {  // 4 FFMA + 4 integer adds (fma pipe + alu pipe co-issue?)
        int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
        int ia = __float_as_int(a);
        for (int i = 0; i < iters; i++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            i0 = (i0 ^ ia) + i; i1 = (i1 ^ ia) + i; i2 = (i2 ^ ia) + i; i3 = (i3 ^ ia) + i;
        }
        x4 = i0 + i1 + i2 + i3;
    } else if (MODE == 4) {  // 8 FFMA with 8 FMNMX (alu pipe) interleaved
        for (int i = 0; i < iters; i++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaxf(x4, x0); x5 = fmaxf(x5, x1); x6 = fminf(x6, x2); x7 = fminf(x7, x3);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int MODE>
static void run_bench(const char* name, float ops_per_iter) {
    float* d;
    int blocks = 148 * 8, threads = 256, iters = 20000;
    CK(cudaMalloc(&d, blocks * threads * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fma_bench<MODE><<<blocks, threads>>>(d, 100, 1.0001f, 0.5f);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    fma_bench<MODE><<<blocks, threads>>>(d, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = (double)blocks * threads * iters * ops_per_iter;
    printf("BENCH %-28s %8.3f ms  %8.2f Tlane-op/s  (%.1f lane-ops/clk/SM @1.9GHz)\n", name, ms, lane_ops / ms * 1e-9,
           lane_ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(d);
}

int main() {
    // ---- ramp texture 4x4x4 -------------------------------------------------------------------
    FILE* f;
    {
        const int n = 4;
        for (int axis = 0; axis < 3; axis++) {
            std::vector<float> h(n * n * n);
            for (int z = 0; z < n; z++)
                for (int y = 0; y < n; y++)
                    for (int x = 0; x < n; x++) h[(z * n + y) * n + x] = axis == 0 ? x : (axis == 1 ? y : z);
            cudaArray_t arr;
            cudaChannelFormatDesc fd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
            CK(cudaMalloc3DArray(&arr, &fd, make_cudaExtent(n, n, n)));
            cudaMemcpy3DParms p = {};
            p.srcPtr = make_cudaPitchedPtr(h.data(), n * 4, n, n);
            p.dstArray = arr;
            p.extent = make_cudaExtent(n, n, n);
            p.kind = cudaMemcpyHostToDevice;
            CK(cudaMemcpy3D(&p));
            cudaResourceDesc rd = {};
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = arr;
            cudaTextureDesc td = {};
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModeLinear;
            td.readMode = cudaReadModeElementType;
            cudaTextureObject_t tex;
            CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
            const int N = 65536;
            float* d;
            CK(cudaMalloc(&d, N * 4));
            ramp_kernel<<<N / 256, 256>>>(tex, d, axis, N);
            std::vector<float> out(N);
            CK(cudaMemcpy(out.data(), d, N * 4, cudaMemcpyDeviceToHost));
            f = fopen(axis == 0 ? "gpurun_out/texprobe_ramp.bin" : "gpurun_out/texprobe_ramp.bin", axis == 0 ? "wb" : "ab");
            fwrite(out.data(), 4, N, f);
            fclose(f);
            // quick summary: distinct values and first few transition points
            int distinct = 1, printed = 0;
            for (int k = 1; k < N; k++)
                if (out[k] != out[k - 1]) {
                    distinct++;
                    if (printed < 4) {
                        printf("axis %d transition at k=%d (frac=%.6f): %.8f -> %.8f\n", axis, k, k / 65536.0, out[k - 1], out[k]);
                        printed++;
                    }
                }
            printf("axis %d: %d distinct values over the unit interval; first=%.8f last=%.8f\n", axis, distinct, out[0], out[N - 1]);
            cudaFree(d);
            cudaDestroyTextureObject(tex);
            cudaFreeArray(arr);
        }
    }
    // ---- random volume -----------------------------------------------------------------------
    {
        const int n = 32, N = 1 << 20;
        std::vector<float> h(n * n * n);
        srand(1234);
        for (auto& v : h) v = (float)rand() / RAND_MAX * 2.0f;
        std::vector<float4> c(N);
        for (int k = 0; k < N; k++) {
            // a band of coordinates reaching past both clamped ends
            c[k].x = -1.0f + (float)rand() / RAND_MAX * (n + 2.0f);
            c[k].y = -1.0f + (float)rand() / RAND_MAX * (n + 2.0f);
            c[k].z = -1.0f + (float)rand() / RAND_MAX * (n + 2.0f);
            c[k].w = 0;
        }
        cudaArray_t arr;
        cudaChannelFormatDesc fd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
        CK(cudaMalloc3DArray(&arr, &fd, make_cudaExtent(n, n, n)));
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr(h.data(), n * 4, n, n);
        p.dstArray = arr;
        p.extent = make_cudaExtent(n, n, n);
        p.kind = cudaMemcpyHostToDevice;
        CK(cudaMemcpy3D(&p));
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arr;
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex;
        CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        float4 *dc, *dout;
        CK(cudaMalloc(&dc, N * 16));
        CK(cudaMalloc(&dout, N * 16));
        CK(cudaMemcpy(dc, c.data(), N * 16, cudaMemcpyHostToDevice));
        rand_kernel<<<N / 256, 256>>>(tex, dc, dout, N);
        std::vector<float4> out(N);
        CK(cudaMemcpy(out.data(), dout, N * 16, cudaMemcpyDeviceToHost));
        f = fopen("gpurun_out/texprobe_vol.bin", "wb");
        fwrite(h.data(), 4, h.size(), f);
        fclose(f);
        f = fopen("gpurun_out/texprobe_rand.bin", "wb");
        fwrite(out.data(), 16, N, f);
        fclose(f);
        printf("wrote %d random samples\n", N);
    }
    // ---- throughput ---------------------------------------------------------------------------
    run_bench<0>("FFMA x8 (scalar)", 8);
    run_bench<1>("FFMA2 x4 (8 fma)", 8);
    run_bench<2>("FADD x8", 8);
    run_bench<3>("FFMA x4 + (LOP+IADD) x4", 8);
    run_bench<4>("FFMA x4 + FMNMX x4", 8);
    return 0;
}
