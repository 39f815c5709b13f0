// Development microbenchmark: sustained rate of tex3D<float> linear fetches (fp32, 3-D, clamp) on B200
// with a DRR-like coherent access pattern (one thread per pixel, 0.1-voxel steps along a ray).
// Also measures the point-sampled u8 fetch rate and a pure-ALU loop of similar length for scale.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int UNROLL>
__global__ void __launch_bounds__(128) march_tex(cudaTextureObject_t tex, float* out, int W, int H, int steps, float dz, float sx, float sy) {
    int u = blockIdx.x * 16 + (threadIdx.x & 15), v = blockIdx.y * 8 + (threadIdx.x >> 4);
    if (u >= W || v >= H) return;
    float x = 16.0f + u * sx, y = 16.0f + v * sy, z = 1.0f;
    float ddx = 0.013f + 1e-5f * u, ddy = 0.021f + 1e-5f * v;
    float acc = 0.f;
    for (int t = 0; t < steps; t += UNROLL) {
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            acc += tex3D<float>(tex, x, y, z);
            x += ddx; y += ddy; z += dz;
        }
    }
    out[v * W + u] = acc;
}

int main() {
    const int n = 320;  // 320^3 fp32 = 131 MB (> L2)
    std::vector<float> h((size_t)n * n * n);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 977) * 1e-3f;
    cudaArray_t arr;
    cudaChannelFormatDesc fd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    CK(cudaMalloc3DArray(&arr, &fd, make_cudaExtent(n, n, n)));
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(h.data(), n * 4, n, n);
    p.dstArray = arr; p.extent = make_cudaExtent(n, n, n); p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&p));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    td.filterMode = cudaFilterModePoint;
    cudaTextureObject_t texp;
    CK(cudaCreateTextureObject(&texp, &rd, &td, nullptr));
    int W = 1536, H = 1536, steps = 2800;
    float* d; CK(cudaMalloc(&d, (size_t)W * H * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid((W + 15) / 16, (H + 7) / 8);
    for (int mode = 0; mode < 2; mode++) {
        cudaTextureObject_t t = mode == 0 ? tex : texp;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            march_tex<4><<<grid, 128>>>(t, d, W, H, steps, 0.1f, 0.1f, 0.1f);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double s = (double)W * H * steps;
            printf("TEXRATE %s rep%d: %.3f ms, %.3e fetch/s, %.3f fetch/clk/SM @1.9GHz\n", mode == 0 ? "linear-f32-3D" : "point-f32-3D", rep, ms,
                   s / (ms * 1e-3), s / (ms * 1e-3) / 148 / 1.9e9);
        }
    }
    return 0;
}
