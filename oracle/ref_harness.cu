// TEST INFRASTRUCTURE -- not part of the product.
//
// Host harness that drives the reference's own GPU projector kernel: the cubins under oracle/_ref/
// are compiled by oracle/Makefile from /root/reference/deepdrr/projector/project_kernel.cu where it
// lies (no reference source is copied into this repository).  The harness only re-creates what the
// reference's Python glue does around the launch:
//   * create_cuda_texture  (deepdrr/projector/projector.py:116-257): 3-D cudaArray, clamp
//     addressing, element read mode, unnormalised coordinates, linear filter for the f32 density
//     and point filter for the u8 labels; array axes (x, y, z) = (i, j, k) after the (0,1,2)->(2,1,0)
//     axis move of projector.py:1468-1470 / 1509.
//   * the 37-argument projectKernel launch (projector.py:718-774; prototype project_kernel.cu:136-181)
//     with block (threads, threads, 1) and grid ceil(W/threads) x ceil(H/threads) (projector.py:759-774).
//   * optionally the reference's per-view host flow (projector.py:802-831 five small H2D uploads,
//     :786-792 two blocking D2H copies + swapaxes copies) so "reference end to end" can be timed.
// Only tests/, bench.py's reference arm and __graft_entry__.smoke() may load this library.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define RH_MAX_VOL 8

struct RefHarness {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    int V = 0, M = 0, max_mesh_hits = 32, mesh_layers = 2;
    int n_added = 0;
    cudaArray_t vol_arr[RH_MAX_VOL] = {}, seg_arr[RH_MAX_VOL] = {};
    cudaTextureObject_t vol_tex[RH_MAX_VOL] = {}, seg_tex[RH_MAX_VOL] = {};
    int shape[RH_MAX_VOL][3] = {};
    // device-side small arrays
    cudaTextureObject_t *d_vol_tex = nullptr, *d_seg_tex = nullptr;
    int *d_priority = nullptr, *d_enabled = nullptr;
    float *d_min[3] = {}, *d_max[3] = {}, *d_vox[3] = {}, *d_src[3] = {};
    float *d_w2i = nullptr, *d_ijk = nullptr;
    float *d_energies = nullptr, *d_pdf = nullptr, *d_mu = nullptr;
    int n_bins = 0;
    float *d_intensity = nullptr, *d_pprob = nullptr, *d_solid = nullptr;
    bool want_solid = false;
    int out_w = 0, out_h = 0;
    // mesh dummies (the kernel reads mesh_sub_layer_valid[j] unconditionally)
    float *d_hit_alpha = nullptr;
    int8_t *d_hit_facing = nullptr, *d_layer_valid = nullptr;
    float *d_additive = nullptr;
    int *d_mesh_mats = nullptr;
    int n_mesh_mats = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
};

static thread_local std::string g_err;

#define RH_CHECK(x)                                                                       \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            g_err = std::string(#x) + ": " + cudaGetErrorString(e_);                      \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)
#define RH_CU(x)                                                                          \
    do {                                                                                  \
        CUresult r_ = (x);                                                                \
        if (r_ != CUDA_SUCCESS) {                                                         \
            const char* s_ = nullptr;                                                     \
            cuGetErrorString(r_, &s_);                                                    \
            g_err = std::string(#x) + ": " + (s_ ? s_ : "?");                             \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

int ref_create(const char* cubin_path, int device, int V, int M, int max_mesh_hits, int mesh_layers, void** out) {
    RH_CHECK(cudaSetDevice(device));
    RH_CHECK(cudaFree(0));
    RefHarness* h = new RefHarness();
    h->V = V; h->M = M; h->max_mesh_hits = max_mesh_hits; h->mesh_layers = mesh_layers;
    RH_CU(cuModuleLoad(&h->mod, cubin_path));
    RH_CU(cuModuleGetFunction(&h->fn, h->mod, "projectKernel"));
    RH_CHECK(cudaMalloc(&h->d_vol_tex, sizeof(cudaTextureObject_t) * V));
    RH_CHECK(cudaMalloc(&h->d_seg_tex, sizeof(cudaTextureObject_t) * V));
    RH_CHECK(cudaMalloc(&h->d_priority, sizeof(int) * V));
    RH_CHECK(cudaMalloc(&h->d_enabled, sizeof(int) * V));
    for (int a = 0; a < 3; a++) {
        RH_CHECK(cudaMalloc(&h->d_min[a], sizeof(float) * V));
        RH_CHECK(cudaMalloc(&h->d_max[a], sizeof(float) * V));
        RH_CHECK(cudaMalloc(&h->d_vox[a], sizeof(float) * V));
        RH_CHECK(cudaMalloc(&h->d_src[a], sizeof(float) * V));
    }
    RH_CHECK(cudaMalloc(&h->d_w2i, sizeof(float) * 9));
    RH_CHECK(cudaMalloc(&h->d_ijk, sizeof(float) * 12 * V));
    RH_CHECK(cudaMalloc(&h->d_layer_valid, mesh_layers));
    RH_CHECK(cudaMemset(h->d_layer_valid, 0, mesh_layers));
    RH_CHECK(cudaMalloc(&h->d_hit_alpha, 16));
    RH_CHECK(cudaMalloc(&h->d_hit_facing, 16));
    RH_CHECK(cudaMalloc(&h->d_additive, 16));
    RH_CHECK(cudaMalloc(&h->d_mesh_mats, 16));
    RH_CHECK(cudaEventCreate(&h->ev0));
    RH_CHECK(cudaEventCreate(&h->ev1));
    *out = h;
    return 0;
}

// density: float32 [ni][nj][nk] (NumPy C order, k fastest); labels: uint8 already remapped to the
// global material index (projector.py:1499-1509).
int ref_add_volume(void* hp, const float* density, const uint8_t* labels, int ni, int nj, int nk, float sx, float sy,
                   float sz) {
    RefHarness* h = (RefHarness*)hp;
    int v = h->n_added;
    if (v >= h->V || v >= RH_MAX_VOL) { g_err = "too many volumes"; return -1; }
    size_t n = (size_t)ni * nj * nk;
    std::vector<float> dt(n);
    std::vector<uint8_t> lt(n);
    // (i, j, k) -> texture memory [k][j][i]
    for (int i = 0; i < ni; i++)
        for (int j = 0; j < nj; j++) {
            const float* dp = density + ((size_t)i * nj + j) * nk;
            const uint8_t* lp = labels + ((size_t)i * nj + j) * nk;
            for (int k = 0; k < nk; k++) {
                size_t o = ((size_t)k * nj + j) * ni + i;
                dt[o] = dp[k];
                lt[o] = lp[k];
            }
        }
    cudaExtent ext = make_cudaExtent(ni, nj, nk);
    cudaChannelFormatDesc fd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    cudaChannelFormatDesc ud = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
    RH_CHECK(cudaMalloc3DArray(&h->vol_arr[v], &fd, ext));
    RH_CHECK(cudaMalloc3DArray(&h->seg_arr[v], &ud, ext));
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(dt.data(), ni * sizeof(float), ni, nj);
    p.dstArray = h->vol_arr[v];
    p.extent = ext;
    p.kind = cudaMemcpyHostToDevice;
    RH_CHECK(cudaMemcpy3D(&p));
    cudaMemcpy3DParms q = {};
    q.srcPtr = make_cudaPitchedPtr(lt.data(), ni, ni, nj);
    q.dstArray = h->seg_arr[v];
    q.extent = ext;
    q.kind = cudaMemcpyHostToDevice;
    RH_CHECK(cudaMemcpy3D(&q));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    rd.res.array.array = h->vol_arr[v];
    td.filterMode = cudaFilterModeLinear;
    RH_CHECK(cudaCreateTextureObject(&h->vol_tex[v], &rd, &td, nullptr));
    rd.res.array.array = h->seg_arr[v];
    td.filterMode = cudaFilterModePoint;
    RH_CHECK(cudaCreateTextureObject(&h->seg_tex[v], &rd, &td, nullptr));
    h->shape[v][0] = ni; h->shape[v][1] = nj; h->shape[v][2] = nk;
    float mn = -0.5f;
    float mx[3] = {ni - 0.5f, nj - 0.5f, nk - 0.5f};
    float sp[3] = {sx, sy, sz};
    for (int a = 0; a < 3; a++) {
        RH_CHECK(cudaMemcpy(h->d_min[a] + v, &mn, 4, cudaMemcpyHostToDevice));
        RH_CHECK(cudaMemcpy(h->d_max[a] + v, &mx[a], 4, cudaMemcpyHostToDevice));
        RH_CHECK(cudaMemcpy(h->d_vox[a] + v, &sp[a], 4, cudaMemcpyHostToDevice));
    }
    h->n_added++;
    RH_CHECK(cudaMemcpy(h->d_vol_tex, h->vol_tex, sizeof(cudaTextureObject_t) * h->n_added, cudaMemcpyHostToDevice));
    RH_CHECK(cudaMemcpy(h->d_seg_tex, h->seg_tex, sizeof(cudaTextureObject_t) * h->n_added, cudaMemcpyHostToDevice));
    return 0;
}

// Mesh inputs of projectKernel (project_kernel.cu:172-177) from host arrays laid out as the reference keeps them:
// hit_alphas / hit_facing [layers][npix][max_hits], layer_valid [layers], additive [layers][n_mats][npix][2].
int ref_set_mesh(void* hp, const float* hit_alphas, const int8_t* hit_facing, const int8_t* layer_valid, const float* additive,
                 const int* mesh_mats, int n_mesh_mats, int npix) {
    RefHarness* h = (RefHarness*)hp;
    size_t nh = (size_t)h->mesh_layers * npix * h->max_mesh_hits, na = (size_t)h->mesh_layers * n_mesh_mats * npix * 2;
    cudaFree(h->d_hit_alpha); cudaFree(h->d_hit_facing); cudaFree(h->d_additive); cudaFree(h->d_mesh_mats);
    RH_CHECK(cudaMalloc(&h->d_hit_alpha, nh * 4 + 16));
    RH_CHECK(cudaMalloc(&h->d_hit_facing, nh + 16));
    RH_CHECK(cudaMalloc(&h->d_additive, na * 4 + 16));
    RH_CHECK(cudaMalloc(&h->d_mesh_mats, n_mesh_mats * 4 + 16));
    RH_CHECK(cudaMemcpy(h->d_hit_alpha, hit_alphas, nh * 4, cudaMemcpyHostToDevice));
    RH_CHECK(cudaMemcpy(h->d_hit_facing, hit_facing, nh, cudaMemcpyHostToDevice));
    RH_CHECK(cudaMemcpy(h->d_layer_valid, layer_valid, h->mesh_layers, cudaMemcpyHostToDevice));
    RH_CHECK(cudaMemcpy(h->d_additive, additive, na * 4, cudaMemcpyHostToDevice));
    RH_CHECK(cudaMemcpy(h->d_mesh_mats, mesh_mats, n_mesh_mats * 4, cudaMemcpyHostToDevice));
    h->n_mesh_mats = n_mesh_mats;
    return 0;
}

int ref_set_spectrum(void* hp, int n_bins, const float* energies, const float* pdf, const float* mu) {
    RefHarness* h = (RefHarness*)hp;
    if (h->d_energies) { cudaFree(h->d_energies); cudaFree(h->d_pdf); cudaFree(h->d_mu); }
    h->n_bins = n_bins;
    RH_CHECK(cudaMalloc(&h->d_energies, 4 * n_bins));
    RH_CHECK(cudaMalloc(&h->d_pdf, 4 * n_bins));
    RH_CHECK(cudaMalloc(&h->d_mu, 4 * (size_t)n_bins * h->M));
    RH_CHECK(cudaMemcpy(h->d_energies, energies, 4 * n_bins, cudaMemcpyHostToDevice));
    RH_CHECK(cudaMemcpy(h->d_pdf, pdf, 4 * n_bins, cudaMemcpyHostToDevice));
    RH_CHECK(cudaMemcpy(h->d_mu, mu, 4 * (size_t)n_bins * h->M, cudaMemcpyHostToDevice));
    return 0;
}

// One view.  Outputs are the kernel's raw buffers, index udx*H + vdx (project_kernel.cu:208), unless
// `transpose` is set, in which case the host-side swapaxes copies of projector.py:786-792 are done
// and the outputs are [H][W].  kernel_ms receives the CUDA-event time of the launch alone.
int ref_project(void* hp, int W, int H, float step, const int* priority, const int* enabled, const float* src_ijk,
                float max_ray_length, const float* w2i, const float* ijk_from_world, int threads, int transpose,
                float* out_intensity, float* out_pprob, float* kernel_ms) {
    RefHarness* h = (RefHarness*)hp;
    int V = h->V;
    size_t npx = (size_t)W * H;
    if (h->out_w != W || h->out_h != H) {
        if (h->d_intensity) { cudaFree(h->d_intensity); cudaFree(h->d_pprob); cudaFree(h->d_solid); }
        RH_CHECK(cudaMalloc(&h->d_intensity, 4 * npx));
        RH_CHECK(cudaMalloc(&h->d_pprob, 4 * npx));
        RH_CHECK(cudaMalloc(&h->d_solid, 4 * npx));
        h->out_w = W; h->out_h = H;
    }
    // projector.py:802-831 -- five small uploads per view (+ priorities/enabled, projector.py:674-675)
    RH_CHECK(cudaMemcpy(h->d_w2i, w2i, 36, cudaMemcpyHostToDevice));
    RH_CHECK(cudaMemcpy(h->d_ijk, ijk_from_world, 48 * V, cudaMemcpyHostToDevice));
    std::vector<float> s(V);
    for (int a = 0; a < 3; a++) {
        for (int v = 0; v < V; v++) s[v] = src_ijk[v * 3 + a];
        RH_CHECK(cudaMemcpy(h->d_src[a], s.data(), 4 * V, cudaMemcpyHostToDevice));
    }
    RH_CHECK(cudaMemcpy(h->d_priority, priority, 4 * V, cudaMemcpyHostToDevice));
    RH_CHECK(cudaMemcpy(h->d_enabled, enabled, 4 * V, cudaMemcpyHostToDevice));

    // project_kernel.cu:213-216: the kernel fills solid_angle when the pointer is non-null (projector.py:752 passes the
    // buffer when collected_energy is set, else 0)
    void* solid = h->want_solid ? (void*)h->d_solid : nullptr;
    int n_mesh_mats = h->n_mesh_mats, off = 0;
    void* args[] = {&h->d_vol_tex, &h->d_seg_tex, &W, &H, &step, &h->d_priority, &h->d_enabled,
                    &h->d_min[0], &h->d_min[1], &h->d_min[2], &h->d_max[0], &h->d_max[1], &h->d_max[2],
                    &h->d_vox[0], &h->d_vox[1], &h->d_vox[2], &h->d_src[0], &h->d_src[1], &h->d_src[2],
                    &max_ray_length, &h->d_w2i, &h->d_ijk, &h->n_bins, &h->d_energies, &h->d_pdf, &h->d_mu,
                    &h->d_intensity, &h->d_pprob, &solid, &h->d_hit_alpha, &h->d_hit_facing, &h->d_layer_valid,
                    &h->d_additive, &h->d_mesh_mats, &n_mesh_mats, &off, &off};
    unsigned bw = (W + threads - 1) / threads, bh = (H + threads - 1) / threads;
    RH_CHECK(cudaEventRecord(h->ev0, 0));
    RH_CU(cuLaunchKernel(h->fn, bw, bh, 1, threads, threads, 1, 0, 0, args, nullptr));
    RH_CHECK(cudaEventRecord(h->ev1, 0));
    if (out_intensity) {
        if (!transpose) {
            RH_CHECK(cudaMemcpy(out_intensity, h->d_intensity, 4 * npx, cudaMemcpyDeviceToHost));
            RH_CHECK(cudaMemcpy(out_pprob, h->d_pprob, 4 * npx, cudaMemcpyDeviceToHost));
        } else {
            std::vector<float> tmp(npx);
            float* outs[2] = {out_intensity, out_pprob};
            float* srcs[2] = {h->d_intensity, h->d_pprob};
            for (int o = 0; o < 2; o++) {
                RH_CHECK(cudaMemcpy(tmp.data(), srcs[o], 4 * npx, cudaMemcpyDeviceToHost));
                for (int u = 0; u < W; u++)
                    for (int v = 0; v < H; v++) outs[o][(size_t)v * W + u] = tmp[(size_t)u * H + v];
            }
        }
    }
    RH_CHECK(cudaEventSynchronize(h->ev1));
    if (kernel_ms) RH_CHECK(cudaEventElapsedTime(kernel_ms, h->ev0, h->ev1));
    RH_CHECK(cudaGetLastError());
    return 0;
}

// kernelTide of the reference (peel_postprocess_kernel.cu:158-177, launch as projector.py:1208-1231) on host
// arrays: ts [n_rays][32] in the peel layout (-exit, +exit, -entry, +entry per pass), facing [n_rays][32] out.
int ref_tide(const char* cubin_path, float* ts, int8_t* facing, int n_rays, float source_to_detector_distance_x2) {
    RH_CHECK(cudaFree(0));
    CUmodule mod;
    CUfunction fn;
    RH_CU(cuModuleLoad(&mod, cubin_path));
    RH_CU(cuModuleGetFunction(&fn, mod, "kernelTide"));
    float* dt; int8_t* df;
    size_t n = (size_t)n_rays * 32;
    RH_CHECK(cudaMalloc(&dt, n * 4));
    RH_CHECK(cudaMalloc(&df, n));
    RH_CHECK(cudaMemcpy(dt, ts, n * 4, cudaMemcpyHostToDevice));
    RH_CHECK(cudaMemset(df, 0, n));
    void* args[] = {&dt, &df, &n_rays, &source_to_detector_distance_x2};
    RH_CU(cuLaunchKernel(fn, 2048, 1, 1, 32, 1, 1, 0, 0, args, nullptr));
    RH_CHECK(cudaDeviceSynchronize());
    RH_CHECK(cudaMemcpy(ts, dt, n * 4, cudaMemcpyDeviceToHost));
    RH_CHECK(cudaMemcpy(facing, df, n, cudaMemcpyDeviceToHost));
    cudaFree(dt); cudaFree(df);
    cuModuleUnload(mod);
    return 0;
}

int ref_destroy(void* hp) {
    RefHarness* h = (RefHarness*)hp;
    if (!h) return 0;
    for (int v = 0; v < h->n_added; v++) {
        cudaDestroyTextureObject(h->vol_tex[v]);
        cudaDestroyTextureObject(h->seg_tex[v]);
        cudaFreeArray(h->vol_arr[v]);
        cudaFreeArray(h->seg_arr[v]);
    }
    cudaFree(h->d_vol_tex); cudaFree(h->d_seg_tex); cudaFree(h->d_priority); cudaFree(h->d_enabled);
    for (int a = 0; a < 3; a++) { cudaFree(h->d_min[a]); cudaFree(h->d_max[a]); cudaFree(h->d_vox[a]); cudaFree(h->d_src[a]); }
    cudaFree(h->d_w2i); cudaFree(h->d_ijk); cudaFree(h->d_energies); cudaFree(h->d_pdf); cudaFree(h->d_mu);
    cudaFree(h->d_intensity); cudaFree(h->d_pprob); cudaFree(h->d_solid); cudaFree(h->d_layer_valid); cudaFree(h->d_hit_alpha);
    cudaFree(h->d_hit_facing); cudaFree(h->d_additive); cudaFree(h->d_mesh_mats);
    cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
    if (h->mod) cuModuleUnload(h->mod);
    delete h;
    return 0;
}

// Ask the next ref_project calls to pass a solid_angle buffer, and read it back: raw kernel layout [W][H]
// (index udx * H + vdx, project_kernel.cu:208) or, with `transpose`, [H][W] like the images.
int ref_want_solid(void* hp, int on) { ((RefHarness*)hp)->want_solid = on != 0; return 0; }

int ref_fetch_solid(void* hp, float* out, int transpose) {
    RefHarness* h = (RefHarness*)hp;
    if (!h->d_solid || !h->want_solid) { g_err = "ref_fetch_solid: no solid-angle buffer (call ref_want_solid and ref_project first)"; return -1; }
    const int W = h->out_w, H = h->out_h;
    const size_t npx = (size_t)W * H;
    if (!transpose) { RH_CHECK(cudaMemcpy(out, h->d_solid, 4 * npx, cudaMemcpyDeviceToHost)); return 0; }
    std::vector<float> tmp(npx);
    RH_CHECK(cudaMemcpy(tmp.data(), h->d_solid, 4 * npx, cudaMemcpyDeviceToHost));
    for (int u = 0; u < W; u++)
        for (int v = 0; v < H; v++) out[(size_t)v * W + u] = tmp[(size_t)u * H + v];
    return 0;
}

// End of a run: nothing of the harness may still be executing when the process exits.
int ref_shutdown() {
    cudaDeviceSynchronize();
    return 0;
}

}  // extern "C"
