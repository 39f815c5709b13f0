// C ABI of libdrr_b200.so (include/drr_b200.h): handle management, volume upload / cell-record
// construction, batch projection.  Host side of the B200-native replacement for the GPU glue in
// /root/reference/deepdrr/projector/projector.py (initialize 1395-1717, project 655-800, free 1719-1764).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "drr_device.cuh"

// kernels / launchers defined in the other translation units
cudaError_t drr_launch_march_single(const MarchParams& P, int grid, cudaStream_t s);
cudaError_t drr_launch_march_general(const MarchParams& P, cudaStream_t s);
cudaError_t drr_launch_march_warp(const MarchParams& P, int n_sm, cudaStream_t s);
int drr_march_single_occupancy(int M);
cudaError_t drr_launch_spectral(const float* area, int n_bins, int M, const float* energies, const float* pdf, const float* mu,
                                size_t npix, int n_views, float* intensity, float* pprob, int n_sm, cudaStream_t s);
cudaError_t drr_launch_noise(float* intensity, const float* pprob, float* scratch, int W, int H, int n_views, float photon_count,
                             unsigned long long seed, cudaStream_t s);
cudaError_t drr_launch_clip(float* img, size_t total, float upper, cudaStream_t s);
cudaError_t drr_launch_neglog(float* img, size_t npix, int n_views, unsigned* minmax, float epsilon, cudaStream_t s, int* const_flag = nullptr);
cudaError_t drr_launch_collected(float* intensity, float* solid, double* view_sum, const ViewDev* views, int W, int H, int n_views,
                                 float photon_count, float pixel_area, cudaStream_t s);

struct MeshPrimDev {
    int tri_begin, tri_end;
    int mat_slot;
    int layer;
    float density;
    int additive, subtractive;
};
cudaError_t drr_launch_mesh_transform(const float* verts_local, const int* prim_of_tri, const float* world_from_mesh, int n_tris, int n_prims,
                                      int n_views, float* verts_world, cudaStream_t s);
cudaError_t drr_launch_mesh_subtractive(const ViewDev* views, const float* source_world, const float* verts_world, const MeshPrimDev* prims,
                                        int n_prims, int n_tris, int layer, int n_layers, int W, int H, int n_views, int max_hits,
                                        float far_limit, float* hit_alphas, int8_t* hit_facing, const int4* tri_box, const int* prim_box,
                                        cudaStream_t s);
cudaError_t drr_launch_mesh_additive(const ViewDev* views, const float* source_world, const float* verts_world, const MeshPrimDev* prims,
                                     int n_prims, int n_tris, int n_layers, int n_mats, int W, int H, int n_views, int max_hits,
                                     const int8_t* layer_valid, const float* hit_alphas, const int8_t* hit_facing, float* additive,
                                     const int4* tri_box, const int* prim_box, cudaStream_t s);
cudaError_t drr_launch_mesh_project(const ViewDev* views, const float* source_world, const float* verts_world, const int* prim_of_tri,
                                    int n_tris, int n_prims, int n_views, int4* tri_box, int* prim_box, cudaStream_t s);
cudaError_t drr_launch_tide_clean(float* ts, int8_t* facing, int n_rays, int n, float far_limit, cudaStream_t s);
cudaError_t drr_launch_march_meshonly(const MarchParams& P, cudaStream_t s);
cudaError_t drr_launch_march_multi(const MarchParams& P, int n_sm, cudaStream_t s);
cudaError_t drr_launch_march_general_list(const MarchParams& P, int grid, cudaStream_t s);
cudaError_t drr_launch_mesh_cover(const ViewDev* views, const float* source_world, const float* verts_world, const MeshPrimDev* prims,
                                  int n_prims, int W, int H, uint8_t* out, const int4* tri_box, const int* prim_box, cudaStream_t s);
cudaError_t drr_launch_mesh_travel_finish(const float* rg, int npix, float* out, cudaStream_t s);

struct ScatterTables {
    int n_mat, n_e;
    float e0, de;
    const float* mfp;
    const float* rita;
    const float* compton;
    const int* nshell;
    const float* inv_rho_nom;
    const float* majorant;
    const int* mat_of_label;
    const float* s0;
};
struct ScatterParams {
    ScatterTables T;
    int V;
    int priority[DRR_MAX_VOLUMES];
    int enabled[DRR_MAX_VOLUMES];
    VolDev vol[DRR_MAX_VOLUMES];
    float ijk[DRR_MAX_VOLUMES][12];
    float p_idx[12];
    float w2i[9];
    float src[3];
    int W, H;
    int n_bins;
    const float* spec_e_keV;
    const float* spec_cdf;
    unsigned long long n_photons, photon_offset, seed;
    unsigned long long* tally;
    double* counters;
};
cudaError_t drr_launch_scatter(const ScatterParams& P, int n_sm, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// upload kernels
// ---------------------------------------------------------------------------------------------
// [ni][nj][nk] (k fastest, NumPy order of Volume.data) -> [nk][nj][ni] (i fastest, texture order;
// the reference does this with cupy.moveaxis, projector.py:1468-1470 / 1509).
template <typename T>
__global__ void transpose_ik_kernel(const T* __restrict__ in, T* __restrict__ out, int ni, int nj, int nk) {
    __shared__ T tile[32][33];
    const int j = blockIdx.z;
    const int i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int i = i0 + r, k = k0 + threadIdx.x;
        if (i < ni && k < nk) tile[r][threadIdx.x] = in[((size_t)i * nj + j) * nk + k];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = k0 + r, i = i0 + threadIdx.x;
        if (i < ni && k < nk) out[((size_t)k * nj + j) * ni + i] = tile[threadIdx.x][r];
    }
}

// HU -> (density, label) on the device.  Same float32 arithmetic as the host path:
// density = max(min(0.001029*HU + 1.03, 0.0005886*HU + 1.03), 0) (vol/volume.py:338-351), labels by the thresholds of
// load_dicom.py:132-143 packed like _format_materials (vol/volume.py:955-992), then remapped to the global index.
__global__ void hu_prepare_kernel(const float* __restrict__ hu, size_t n, int l_air, int l_soft, int l_bone, float* __restrict__ dens,
                                  uint8_t* __restrict__ lab) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float h = hu[i];
    const float a = __fadd_rn(__fmul_rn(0.001029f, h), 1.030f), b = __fadd_rn(__fmul_rn(0.0005886f, h), 1.03f);
    dens[i] = fmaxf(fminf(a, b), 0.0f);
    int l = l_air;                        // unlabeled voxels (NaN) stay at id 0 = the first material, air
    if (-800.0f < h && h <= 350.0f) l = l_soft;
    if (350.0f < h) l = l_bone;
    lab[i] = (uint8_t)l;
}

// One thread per cell base (bi, bj, bk) in [-2, n-2]^3: gathers the 8 clamped corner texels / labels
// and stores the filter-coefficient record (see hw_trilinear_cell) and the label record.
__global__ void build_cells_kernel(const float* __restrict__ dens, const uint8_t* __restrict__ lab, int ni, int nj, int nk,
                                   float4* __restrict__ cellc, uint2* __restrict__ celll, uint8_t* __restrict__ cellcode) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    const int cj = blockIdx.y, ck = blockIdx.z;
    if (ci > ni) return;
    const int bi = ci - 2, bj = cj - 2, bk = ck - 2;
    float T[2][2][2];  // [z][x][y]
    unsigned lx = 0, ly = 0;
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int a = 0; a < 2; a++) {
                int i = max(0, min(bi + a, ni - 1)), j = max(0, min(bj + b, nj - 1)), k = max(0, min(bk + c, nk - 1));
                size_t o = ((size_t)k * nj + j) * ni + i;
                T[c][a][b] = dens[o];
                unsigned l = lab[o];
                if (c) ly |= l << (8 * (a + 2 * b)); else lx |= l << (8 * (a + 2 * b));
            }
    const size_t cell = ((size_t)ck * (nj + 1) + cj) * (ni + 1) + ci;
    const float s = 1.0f / 256.0f;
    // per z-slice c: (T01, T10 - T01, T00 - T01, T11 - T10) / 256; the two slices are stored interleaved,
    // A = (c0.x, c1.x, c0.y, c1.y), B = (c0.z, c1.z, c0.w, c1.w), which is the operand order of the f32x2 form
    float4 cs[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        float t01 = T[c][0][1], t10 = T[c][1][0], t00 = T[c][0][0], t11 = T[c][1][1];
        cs[c] = make_float4(t01 * s, (t10 - t01) * s, (t00 - t01) * s, (t11 - t10) * s);
    }
    if (cellc != nullptr) {  // absent when the records did not fit in device memory (texture-unit sampler only)
        cellc[2 * cell] = make_float4(cs[0].x, cs[1].x, cs[0].y, cs[1].y);
        cellc[2 * cell + 1] = make_float4(cs[0].z, cs[1].z, cs[0].w, cs[1].w);
    }
    celll[cell] = make_uint2(lx, ly);
    const unsigned l0 = lx & 0xFF;
    const bool uniform = (lx == ly) && (lx == l0 * 0x01010101u);
    // clamped cells (base < 0) and the first voxel layer (base == 0) take the integer path, see hw_trilinear_cell
    const bool interior = bi >= 1 && bj >= 1 && bk >= 1;
    cellcode[cell] = (uniform && interior) ? (uint8_t)l0 : (uint8_t)0xFF;
}

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct VolHost {
    float* dens = nullptr;
    uint8_t* lab = nullptr;
    float4* cellc = nullptr;
    uint2* celll = nullptr;
    uint8_t* cellcode = nullptr;
    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
    int ni = 0, nj = 0, nk = 0;
};

struct drr_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t evp[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // copy / compute pipeline of drr_project
    int* d_const_flag = nullptr;
    int h_const_flag = 0;
    int pipeline = 1;  // DRR_TUNE_PIPELINE: host-bound batches are projected in two pieces, the copy of the first under the march of the
                       // second; the value is the number of views in the second piece (0 = one piece)
    std::string err;
    // spectrum
    int n_bins = 0, M = 0;
    float *d_energies = nullptr, *d_pdf = nullptr, *d_mu = nullptr;
    // volumes
    std::vector<VolHost> vols;
    int priority[DRR_MAX_VOLUMES];
    int enabled[DRR_MAX_VOLUMES];
    bool priorities_set = false;
    // march options
    float step = 0.1f;
    int attenuate_outside = 0, air_index = 0, sampler = DRR_SAMPLER_HYBRID;
    int tex_eighths = 4;
    int variant = 0;  // 0: warp-cooperative march (default), 1: per-ray register-cell march
    int lane_quads = 2;  // DRR_TUNE_LANE_QUADS: 0 / 1, 2 = the library's choice (1)
    int rays_per_lane = 0;  // DRR_TUNE_RAYS_PER_LANE: 1 / 2, 0 = by ray spacing
    // mesh buffers (device pointers, possibly owned)
    int mesh_layers = 0, max_hits = 0, n_mesh_mats = 0;
    const float* hit_alphas = nullptr;
    const int8_t* hit_facing = nullptr;
    const int8_t* layer_valid = nullptr;
    const float* additive = nullptr;
    const int* mesh_mats = nullptr;
    std::vector<void*> mesh_owned;
    // meshes traced by this library (drr_set_meshes / drr_set_mesh_poses)
    int n_prims = 0, n_tris = 0, own_layers = 0, own_max_hits = 0, own_n_mats = 0;
    std::vector<MeshPrimDev> h_prims;
    std::vector<int8_t> h_layer_valid;
    float* d_verts_local = nullptr; int* d_prim_of_tri = nullptr; MeshPrimDev* d_prims = nullptr; int* d_own_mesh_mats = nullptr;
    int8_t* d_own_layer_valid = nullptr;
    std::vector<float> h_world_from_mesh, h_source_world;
    int pose_views = 0; float far_limit = 0.0f;
    // scatter
    std::vector<float> h_energies, h_pdf;
    ScatterTables sc = {};
    bool sc_ready = false;
    float* d_sc_cdf = nullptr; unsigned long long* d_sc_tally = nullptr; double* d_sc_counters = nullptr; size_t sc_tally_cap = 0;
    std::vector<void*> sc_owned;
    float *d_world_from_mesh = nullptr, *d_source_world = nullptr, *d_verts_world = nullptr, *d_own_hit_alphas = nullptr, *d_own_additive = nullptr;
    int8_t* d_own_hit_facing = nullptr;
    size_t wfm_cap = 0, srcw_cap = 0, vw_cap = 0, oha_cap = 0, ohf_cap = 0, oadd_cap = 0;
    unsigned int* d_worklist = nullptr; size_t worklist_cap = 0;
    int4* d_tri_box = nullptr; int* d_prim_box = nullptr; size_t tribox_cap = 0, primbox_cap = 0;
    // per-batch scratch (grown on demand)
    ViewDev* d_views = nullptr; ViewDev* h_views = nullptr; int views_cap = 0;
    float *d_area = nullptr, *d_intensity = nullptr, *d_pprob = nullptr, *d_scratch = nullptr;
    size_t area_cap = 0, int_cap = 0, pp_cap = 0, scratch_cap = 0;
    unsigned* d_minmax = nullptr; double* d_viewsum = nullptr; int minmax_cap = 0;
    unsigned long long* d_samples = nullptr;
    unsigned int* d_tile_counter = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float last_ms[3] = {0, 0, 0};
    unsigned long long last_samples[2] = {0, 0}, launches = 0;
};

static thread_local std::string g_create_err;

static int fail(drr_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_err = buf;
    return code;
}

// 3x3 inverse (double) of world_from_index: maps a world vector from the source to its homogeneous pixel.
static void invert3(const float* m, float* out) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    const double r = det != 0.0 ? 1.0 / det : 0.0;
    out[0] = (float)(A * r); out[1] = (float)(-(b * i - c * h) * r); out[2] = (float)((b * f - c * e) * r);
    out[3] = (float)(B * r); out[4] = (float)((a * i - c * g) * r);  out[5] = (float)(-(a * f - c * d) * r);
    out[6] = (float)(C * r); out[7] = (float)(-(a * h - b * g) * r); out[8] = (float)((a * e - b * d) * r);
}

extern "C" {
static int ensure(drr_ctx* c, void** p, size_t* cap, size_t bytes);
}

#define CU(c, x)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess)                                                                       \
            return fail(c, e_ == cudaErrorMemoryAllocation ? DRR_E_NOMEM : DRR_E_CUDA, "%s: %s", #x, \
                        cudaGetErrorString(e_));                                                     \
    } while (0)

static void free_volume(VolHost& v) {
    if (v.tex) cudaDestroyTextureObject(v.tex);
    if (v.arr) cudaFreeArray(v.arr);
    cudaFree(v.dens); cudaFree(v.lab); cudaFree(v.cellc); cudaFree(v.celll); cudaFree(v.cellcode);
    v = VolHost();
}

extern "C" {

const char* drr_version(void) { return "drr_b200 0.1.0 sm_100a"; }

const char* drr_last_error(const drr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int drr_create(int device_id, drr_ctx** out) {
    if (!out) return fail(nullptr, DRR_E_INVALID, "drr_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, DRR_E_CUDA, "drr_create: no CUDA device (%s); libdrr_b200 has no CPU fallback",
                    cudaGetErrorString(e));
    if (device_id < 0 || device_id >= n) return fail(nullptr, DRR_E_INVALID, "drr_create: device %d out of range [0,%d)", device_id, n);
    CU(nullptr, cudaSetDevice(device_id));
    drr_ctx* c = new (std::nothrow) drr_ctx();
    if (!c) return fail(nullptr, DRR_E_NOMEM, "drr_create: out of host memory");
    c->device = device_id;
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device_id));
    c->n_sm = prop.multiProcessorCount;
    CU(nullptr, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU(nullptr, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 5; i++) CU(nullptr, cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 7; i++) CU(nullptr, cudaEventCreate(&c->evp[i]));
    CU(nullptr, cudaMalloc(&c->d_const_flag, sizeof(int)));
    CU(nullptr, cudaMalloc(&c->d_samples, 2 * sizeof(unsigned long long)));  // [0] march steps, [1] steps inside a volume window
    CU(nullptr, cudaMalloc(&c->d_tile_counter, sizeof(unsigned int)));
    for (int i = 0; i < DRR_MAX_VOLUMES; i++) { c->priority[i] = 0; c->enabled[i] = 1; }
    *out = c;
    return DRR_OK;
}

int drr_clear_volumes(drr_ctx* c) {
    if (!c) return DRR_E_INVALID;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& v : c->vols) free_volume(v);
    c->vols.clear();
    c->priorities_set = false;
    return DRR_OK;
}

int drr_destroy(drr_ctx* c) {
    if (!c) return DRR_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    drr_clear_volumes(c);
    for (void* p : c->mesh_owned) cudaFree(p);
    for (void* p : c->sc_owned) cudaFree(p);
    cudaFree(c->d_sc_cdf); cudaFree(c->d_sc_tally); cudaFree(c->d_sc_counters);
    cudaFree(c->d_verts_local); cudaFree(c->d_prim_of_tri); cudaFree(c->d_prims); cudaFree(c->d_own_mesh_mats); cudaFree(c->d_own_layer_valid);
    cudaFree(c->d_worklist); cudaFree(c->d_tri_box); cudaFree(c->d_prim_box);
    cudaFree(c->d_world_from_mesh); cudaFree(c->d_source_world); cudaFree(c->d_verts_world); cudaFree(c->d_own_hit_alphas);
    cudaFree(c->d_own_additive); cudaFree(c->d_own_hit_facing);
    cudaFree(c->d_energies); cudaFree(c->d_pdf); cudaFree(c->d_mu);
    cudaFree(c->d_views); cudaFreeHost(c->h_views);
    cudaFree(c->d_area); cudaFree(c->d_intensity); cudaFree(c->d_pprob); cudaFree(c->d_scratch);
    cudaFree(c->d_minmax); cudaFree(c->d_viewsum); cudaFree(c->d_samples); cudaFree(c->d_tile_counter);
    for (int i = 0; i < 5; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 7; i++) if (c->evp[i]) cudaEventDestroy(c->evp[i]);
    cudaFree(c->d_const_flag);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return DRR_OK;
}

int drr_set_stream(drr_ctx* c, void* s) {
    if (!c) return DRR_E_INVALID;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return DRR_OK;
}

int drr_synchronize(drr_ctx* c) {
    if (!c) return DRR_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    return DRR_OK;
}

int drr_set_spectrum(drr_ctx* c, int n_bins, int M, const float* energies, const float* pdf, const float* mu) {
    if (!c) return DRR_E_INVALID;
    if (n_bins <= 0 || M <= 0 || M > DRR_MAX_MATERIALS || !energies || !pdf || !mu)
        return fail(c, DRR_E_INVALID, "drr_set_spectrum: need n_bins > 0, 1 <= n_materials <= %d and non-NULL tables", DRR_MAX_MATERIALS);
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_energies); cudaFree(c->d_pdf); cudaFree(c->d_mu);
    c->d_energies = c->d_pdf = c->d_mu = nullptr;
    CU(c, cudaMalloc(&c->d_energies, sizeof(float) * n_bins));
    CU(c, cudaMalloc(&c->d_pdf, sizeof(float) * n_bins));
    CU(c, cudaMalloc(&c->d_mu, sizeof(float) * (size_t)n_bins * M));
    CU(c, cudaMemcpy(c->d_energies, energies, sizeof(float) * n_bins, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->d_pdf, pdf, sizeof(float) * n_bins, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->d_mu, mu, sizeof(float) * (size_t)n_bins * M, cudaMemcpyHostToDevice));
    c->n_bins = n_bins;
    c->M = M;
    c->h_energies.assign(energies, energies + n_bins);
    c->h_pdf.assign(pdf, pdf + n_bins);
    return DRR_OK;
}

static int add_volume_impl(drr_ctx* c, const float* density, const uint8_t* labels, const float* hu, const int* hu_labels, int ni, int nj, int nk,
                           int mem_kind, unsigned flags, int* vol_id) {
    if (!c) return DRR_E_INVALID;
    if ((!hu && (!density || !labels)) || ni <= 0 || nj <= 0 || nk <= 0) return fail(c, DRR_E_INVALID, "drr_add_volume: bad arguments");
    if ((int)c->vols.size() >= DRR_MAX_VOLUMES) return fail(c, DRR_E_INVALID, "drr_add_volume: at most %d volumes", DRR_MAX_VOLUMES);
    if (ni > 16384 || nj > 16384 || nk > 16384) return fail(c, DRR_E_INVALID, "drr_add_volume: dimension above 16384");
    if ((size_t)(ni + 1) * (nj + 1) * (nk + 1) >= ((size_t)1 << 31))  // 8.6 GB of density alone; the march kernels index cells with 32 bits
        return fail(c, DRR_E_INVALID, "drr_add_volume: more than 2^31 voxel cells");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t n = (size_t)ni * nj * nk;
    VolHost v;
    v.ni = ni; v.nj = nj; v.nk = nk;
    float* d_in = nullptr;
    uint8_t* l_in = nullptr;
    float* hu_dev = nullptr;
    const bool own_in = (mem_kind == DRR_MEM_HOST) || hu;
    auto cleanup = [&]() { free_volume(v); if (own_in) { cudaFree(d_in); cudaFree(l_in); } if (hu && mem_kind == DRR_MEM_HOST) cudaFree(hu_dev); };
#define CUV(x)                                                                                                     \
    do {                                                                                                           \
        cudaError_t e_ = (x);                                                                                      \
        if (e_ != cudaSuccess) {                                                                                   \
            cleanup();                                                                                             \
            return fail(c, e_ == cudaErrorMemoryAllocation ? DRR_E_NOMEM : DRR_E_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); \
        }                                                                                                          \
    } while (0)
    if (hu) {
        CUV(cudaMalloc(&d_in, n * sizeof(float)));
        CUV(cudaMalloc(&l_in, n));
        hu_dev = const_cast<float*>(hu);
        if (mem_kind == DRR_MEM_HOST) {
            CUV(cudaMalloc(&hu_dev, n * sizeof(float)));
            CUV(cudaMemcpyAsync(hu_dev, hu, n * sizeof(float), cudaMemcpyHostToDevice, s));
        }
        hu_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(hu_dev, n, hu_labels[0], hu_labels[1], hu_labels[2], d_in, l_in);
        c->launches += 1;
        CUV(cudaGetLastError());
    } else if (mem_kind == DRR_MEM_HOST) {
        CUV(cudaMalloc(&d_in, n * sizeof(float)));
        CUV(cudaMalloc(&l_in, n));
        CUV(cudaMemcpyAsync(d_in, density, n * sizeof(float), cudaMemcpyHostToDevice, s));
        CUV(cudaMemcpyAsync(l_in, labels, n, cudaMemcpyHostToDevice, s));
    } else {
        d_in = const_cast<float*>(density);
        l_in = const_cast<uint8_t*>(labels);
    }
    CUV(cudaMalloc(&v.dens, n * sizeof(float)));
    CUV(cudaMalloc(&v.lab, n));
    dim3 tb(32, 8), tg((nk + 31) / 32, (ni + 31) / 32, nj);
    transpose_ik_kernel<float><<<tg, tb, 0, s>>>(d_in, v.dens, ni, nj, nk);
    transpose_ik_kernel<uint8_t><<<tg, tb, 0, s>>>(l_in, v.lab, ni, nj, nk);
    c->launches += 2;
    CUV(cudaGetLastError());
    if (!(flags & 1u)) {
        const size_t ncell = (size_t)(ni + 1) * (nj + 1) * (nk + 1);
        CUV(cudaMalloc(&v.celll, ncell * sizeof(uint2)));
        CUV(cudaMalloc(&v.cellcode, ncell));
        // The 32 B / cell coefficient records only serve the FMA-pipe sampler.  If they do not fit next to a texture, go
        // without them: such a volume is sampled by the texture unit alone (drr_project picks the sampler per scene).
        {
            size_t free_b = 0, total_b = 0;
            const size_t want = ncell * 2 * sizeof(float4), tex_bytes = (flags & 2u) ? 0 : n * sizeof(float);
            cudaError_t info = cudaMemGetInfo(&free_b, &total_b);
            if ((flags & 4u) || (info == cudaSuccess && !(flags & 2u) && want + tex_bytes + (size_t)(1u << 28) > free_b)) {
                v.cellc = nullptr;
            } else {
                const cudaError_t got = cudaMalloc(&v.cellc, want);  // (not named e_: CUV declares its own e_)
                if (got == cudaErrorMemoryAllocation && !(flags & 2u)) { cudaGetLastError(); v.cellc = nullptr; }
                else CUV(got);
            }
        }
        dim3 cg((ni + 1 + 127) / 128, nj + 1, nk + 1);
        build_cells_kernel<<<cg, 128, 0, s>>>(v.dens, v.lab, ni, nj, nk, v.cellc, v.celll, v.cellcode);
        c->launches += 1;
        CUV(cudaGetLastError());
    }
    if (!(flags & 2u)) {
        cudaChannelFormatDesc fd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
        CUV(cudaMalloc3DArray(&v.arr, &fd, make_cudaExtent(ni, nj, nk)));
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr(v.dens, (size_t)ni * sizeof(float), ni, nj);
        p.dstArray = v.arr;
        p.extent = make_cudaExtent(ni, nj, nk);
        p.kind = cudaMemcpyDeviceToDevice;
        CUV(cudaMemcpy3DAsync(&p, s));
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = v.arr;
        cudaTextureDesc td = {};  // projector.py:189-235: clamp, linear, element type, unnormalised coordinates
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        CUV(cudaCreateTextureObject(&v.tex, &rd, &td, nullptr));
    }
    CUV(cudaStreamSynchronize(s));
    if (own_in) { cudaFree(d_in); cudaFree(l_in); }
    if (hu && mem_kind == DRR_MEM_HOST) cudaFree(hu_dev);
#undef CUV
    c->vols.push_back(v);
    if (vol_id) *vol_id = (int)c->vols.size() - 1;
    return DRR_OK;
}

int drr_add_volume(drr_ctx* c, const float* density, const uint8_t* labels, int ni, int nj, int nk, int mem_kind, unsigned flags,
                   int* vol_id) {
    return add_volume_impl(c, density, labels, nullptr, nullptr, ni, nj, nk, mem_kind, flags, vol_id);
}

int drr_add_volume_hu(drr_ctx* c, const float* hu, int ni, int nj, int nk, int mem_kind, const int* labels_air_soft_bone, unsigned flags,
                      int* vol_id) {
    if (!c) return DRR_E_INVALID;
    if (!hu || !labels_air_soft_bone) return fail(c, DRR_E_INVALID, "drr_add_volume_hu: bad arguments");
    for (int k = 0; k < 3; k++)
        if (labels_air_soft_bone[k] < 0 || labels_air_soft_bone[k] >= DRR_MAX_MATERIALS)
            return fail(c, DRR_E_INVALID, "drr_add_volume_hu: material index out of range");
    return add_volume_impl(c, nullptr, nullptr, hu, labels_air_soft_bone, ni, nj, nk, mem_kind, flags, vol_id);
}

int drr_set_priorities(drr_ctx* c, const int* priority, const int* enabled, int n) {
    if (!c) return DRR_E_INVALID;
    if (n < 0 || n > DRR_MAX_VOLUMES) return fail(c, DRR_E_INVALID, "drr_set_priorities: bad volume count %d", n);
    for (int i = 0; i < n; i++) {
        if (priority) c->priority[i] = priority[i];
        if (enabled) c->enabled[i] = enabled[i];
    }
    if (priority) c->priorities_set = true;
    return DRR_OK;
}

int drr_set_march(drr_ctx* c, float step, int attenuate_outside, int air_index, int sampler) {
    if (!c) return DRR_E_INVALID;
    if (!(step > 0.0f)) return fail(c, DRR_E_INVALID, "drr_set_march: step must be positive");
    if (sampler < 0 || sampler > 2) return fail(c, DRR_E_INVALID, "drr_set_march: unknown sampler %d", sampler);
    c->step = step;
    c->attenuate_outside = attenuate_outside ? 1 : 0;
    c->air_index = air_index;
    c->sampler = sampler;
    return DRR_OK;
}

int drr_set_tuning(drr_ctx* c, int key, int value) {  // tuning knobs (results do not depend on them)
    if (!c) return DRR_E_INVALID;
    if (key == DRR_TUNE_TEX_EIGHTHS && value >= 0 && value <= 8) { c->tex_eighths = value; return DRR_OK; }
    if (key == DRR_TUNE_KERNEL_VARIANT && (value == 0 || value == 1)) { c->variant = value; return DRR_OK; }
    if (key == DRR_TUNE_PIPELINE && value >= 0 && value <= 64) { c->pipeline = value; return DRR_OK; }
    if (key == DRR_TUNE_LANE_QUADS && value >= 0 && value <= 2) { c->lane_quads = value; return DRR_OK; }
    if (key == DRR_TUNE_RAYS_PER_LANE && value >= 0 && value <= 2) { c->rays_per_lane = value; return DRR_OK; }
    return fail(c, DRR_E_INVALID, "drr_set_tuning: bad key/value %d/%d", key, value);
}

int drr_set_mesh_buffers(drr_ctx* c, int layers, int max_hits, const float* hit_alphas, const int8_t* hit_facing,
                         const int8_t* layer_valid, const float* additive, const int* mesh_mats, int n_mesh_mats, int mem_kind) {
    if (!c) return DRR_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    for (void* p : c->mesh_owned) cudaFree(p);
    c->mesh_owned.clear();
    c->hit_alphas = nullptr; c->hit_facing = nullptr; c->layer_valid = nullptr; c->additive = nullptr; c->mesh_mats = nullptr;
    c->mesh_layers = 0; c->max_hits = 0; c->n_mesh_mats = 0;
    if (!layer_valid && !additive) return DRR_OK;
    if (layers <= 0 || layers > 4) return fail(c, DRR_E_INVALID, "drr_set_mesh_buffers: 1..4 mesh layers supported");
    if (mem_kind != DRR_MEM_DEVICE)
        return fail(c, DRR_E_INVALID, "drr_set_mesh_buffers: pass device pointers (per-batch sizes are only known to the caller)");
    c->mesh_layers = layers; c->max_hits = max_hits; c->n_mesh_mats = n_mesh_mats;
    c->hit_alphas = hit_alphas; c->hit_facing = hit_facing; c->layer_valid = layer_valid;
    c->additive = additive; c->mesh_mats = mesh_mats;
    return DRR_OK;
}

static bool h_has_cells(const drr_ctx* c) { return !c->vols.empty() && c->vols[0].cellcode != nullptr; }

int drr_set_meshes(drr_ctx* c, int n_prims, const int* tri_offsets, const float* vertices, const int* material, const float* density,
                   const uint8_t* flags, const int* layer, int mesh_layers, int max_mesh_hits) {
    if (!c) return DRR_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_verts_local); cudaFree(c->d_prim_of_tri); cudaFree(c->d_prims); cudaFree(c->d_own_mesh_mats); cudaFree(c->d_own_layer_valid);
    c->d_verts_local = nullptr; c->d_prim_of_tri = nullptr; c->d_prims = nullptr; c->d_own_mesh_mats = nullptr; c->d_own_layer_valid = nullptr;
    c->n_prims = 0; c->n_tris = 0; c->h_prims.clear(); c->pose_views = 0;
    if (n_prims == 0) return DRR_OK;
    if (n_prims < 0 || !tri_offsets || !vertices || !material || !density || !flags || !layer)
        return fail(c, DRR_E_INVALID, "drr_set_meshes: bad arguments");
    if (mesh_layers < 1 || mesh_layers > 4) return fail(c, DRR_E_INVALID, "drr_set_meshes: 1..4 mesh layers supported");
    if (max_mesh_hits < 4 || max_mesh_hits % 4 != 0 || max_mesh_hits > 128)
        return fail(c, DRR_E_INVALID, "drr_set_meshes: max_mesh_hits must be a multiple of 4 in [4, 128]");  // projector.py:543-545
    const int n_tris = tri_offsets[n_prims];
    // material slots: sorted unique global material indices (projector.py:1548-1557)
    std::vector<int> mats;
    for (int p = 0; p < n_prims; p++) {
        if (material[p] < 0 || material[p] >= DRR_MAX_MATERIALS) return fail(c, DRR_E_INVALID, "drr_set_meshes: material index out of range");
        bool seen = false;
        for (int m : mats) seen = seen || m == material[p];
        if (!seen) mats.push_back(material[p]);
    }
    for (size_t i = 0; i < mats.size(); i++)
        for (size_t j = i + 1; j < mats.size(); j++)
            if (mats[j] < mats[i]) { int t = mats[i]; mats[i] = mats[j]; mats[j] = t; }
    c->h_prims.resize(n_prims);
    c->h_layer_valid.assign(mesh_layers, 0);
    std::vector<int> prim_of_tri(n_tris);
    for (int p = 0; p < n_prims; p++) {
        MeshPrimDev& d = c->h_prims[p];
        d.tri_begin = tri_offsets[p]; d.tri_end = tri_offsets[p + 1];
        d.layer = layer[p]; d.density = density[p];
        d.additive = flags[p] & 1; d.subtractive = (flags[p] >> 1) & 1;
        d.mat_slot = 0;
        for (size_t i = 0; i < mats.size(); i++) if (mats[i] == material[p]) d.mat_slot = (int)i;
        if (d.subtractive && d.layer >= 0 && d.layer < mesh_layers) c->h_layer_valid[d.layer] = 1;  // projector.py:1127-1142
        for (int t = d.tri_begin; t < d.tri_end; t++) prim_of_tri[t] = p;
    }
    CU(c, cudaMalloc(&c->d_verts_local, sizeof(float) * 9 * (size_t)n_tris));
    CU(c, cudaMalloc(&c->d_prim_of_tri, sizeof(int) * (size_t)n_tris));
    CU(c, cudaMalloc(&c->d_prims, sizeof(MeshPrimDev) * n_prims));
    CU(c, cudaMalloc(&c->d_own_mesh_mats, sizeof(int) * mats.size()));
    CU(c, cudaMalloc(&c->d_own_layer_valid, mesh_layers));
    CU(c, cudaMemcpy(c->d_verts_local, vertices, sizeof(float) * 9 * (size_t)n_tris, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->d_prim_of_tri, prim_of_tri.data(), sizeof(int) * (size_t)n_tris, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->d_prims, c->h_prims.data(), sizeof(MeshPrimDev) * n_prims, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->d_own_mesh_mats, mats.data(), sizeof(int) * mats.size(), cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->d_own_layer_valid, c->h_layer_valid.data(), mesh_layers, cudaMemcpyHostToDevice));
    c->n_prims = n_prims; c->n_tris = n_tris; c->own_layers = mesh_layers; c->own_max_hits = max_mesh_hits; c->own_n_mats = (int)mats.size();
    return DRR_OK;
}

int drr_set_mesh_poses(drr_ctx* c, int n_views, const float* world_from_mesh, const float* source_world, float far_limit) {
    if (!c) return DRR_E_INVALID;
    if (c->n_prims == 0) return fail(c, DRR_E_STATE, "drr_set_mesh_poses: call drr_set_meshes first");
    if (n_views <= 0 || !world_from_mesh || !source_world) return fail(c, DRR_E_INVALID, "drr_set_mesh_poses: bad arguments");
    c->h_world_from_mesh.assign(world_from_mesh, world_from_mesh + (size_t)n_views * c->n_prims * 12);
    c->h_source_world.assign(source_world, source_world + (size_t)n_views * 3);
    c->pose_views = n_views;
    c->far_limit = far_limit;
    return DRR_OK;
}

int drr_set_scatter_tables(drr_ctx* c, int n_mat, int n_e, float e0, float de, const float* mfp, const float* rita, const float* compton,
                           const int* nshell, const float* rho_nom, const int* mat_of_label, const float* rho_max_of_label) {
    if (!c) return DRR_E_INVALID;
    if (c->M == 0) return fail(c, DRR_E_STATE, "drr_set_scatter_tables: call drr_set_spectrum first");
    if (n_mat <= 0 || n_e < 2 || !mfp || !rita || !compton || !nshell || !rho_nom || !mat_of_label || !rho_max_of_label)
        return fail(c, DRR_E_INVALID, "drr_set_scatter_tables: bad arguments");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    for (void* p : c->sc_owned) cudaFree(p);
    c->sc_owned.clear();
    c->sc_ready = false;
    std::vector<float> inv_rho(n_mat), maj(n_e, 0.0f);
    for (int m = 0; m < n_mat; m++) inv_rho[m] = 1.0f / rho_nom[m];
    for (int l = 0; l < c->M; l++) {
        int m = mat_of_label[l];
        if (m < 0 || m >= n_mat) return fail(c, DRR_E_INVALID, "drr_set_scatter_tables: material %d has no MC table", l);
        for (int e = 0; e < n_e; e++) {
            float mu = rho_max_of_label[l] * inv_rho[m] / mfp[((size_t)m * n_e + e) * 5 + 3];
            if (mu > maj[e]) maj[e] = mu;
        }
    }
    for (int e = 0; e < n_e; e++) if (!(maj[e] > 0.0f)) maj[e] = 1e-6f;
    auto up = [&](const void* src, size_t bytes, const void** dst) -> int {
        void* d = nullptr;
        CU(c, cudaMalloc(&d, bytes));
        c->sc_owned.push_back(d);
        CU(c, cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
        *dst = d;
        return DRR_OK;
    };
    int rc;
    const void* p;
    if ((rc = up(mfp, sizeof(float) * 5 * (size_t)n_mat * n_e, &p))) return rc; c->sc.mfp = (const float*)p;
    if ((rc = up(rita, sizeof(float) * 4 * 128 * (size_t)n_mat, &p))) return rc; c->sc.rita = (const float*)p;
    if ((rc = up(compton, sizeof(float) * 3 * 30 * (size_t)n_mat, &p))) return rc; c->sc.compton = (const float*)p;
    if ((rc = up(nshell, sizeof(int) * n_mat, &p))) return rc; c->sc.nshell = (const int*)p;
    if ((rc = up(inv_rho.data(), sizeof(float) * n_mat, &p))) return rc; c->sc.inv_rho_nom = (const float*)p;
    if ((rc = up(maj.data(), sizeof(float) * n_e, &p))) return rc; c->sc.majorant = (const float*)p;
    if ((rc = up(mat_of_label, sizeof(int) * c->M, &p))) return rc; c->sc.mat_of_label = (const int*)p;
    {   // S(E, theta = pi) of the impulse-approximation Compton model per table material on the energy grid (see sample_compton)
        std::vector<float> s0((size_t)n_mat * n_e);
        const double REV = 510998.918, D2 = 1.4142135623731, D1 = 0.70710678118655;
        for (int m = 0; m < n_mat; m++)
            for (int e = 0; e < n_e; e++) {
                const double E = (double)e0 + (double)de * e;
                double sum = 0.0;
                for (int i = 0; i < nshell[m]; i++) {
                    const double f = compton[((size_t)m * 30 + i) * 3], U = compton[((size_t)m * 30 + i) * 3 + 1], J = compton[((size_t)m * 30 + i) * 3 + 2];
                    if (!(U < E)) continue;
                    const double aux = E * (E - U) * 2.0;
                    const double pz = J * (aux - REV * U) / (REV * sqrt(aux + aux + U * U));
                    const double q = pz > 0.0 ? D1 + D2 * pz : D1 - D2 * pz;
                    const double h = 0.5 * exp(0.5 - q * q);
                    sum += f * (pz > 0.0 ? 1.0 - h : h);
                }
                s0[(size_t)m * n_e + e] = (float)(sum * 1.001);
            }
        if ((rc = up(s0.data(), sizeof(float) * s0.size(), &p))) return rc; c->sc.s0 = (const float*)p;
    }
    c->sc.n_mat = n_mat; c->sc.n_e = n_e; c->sc.e0 = e0; c->sc.de = de;
    // spectrum CDF of max(pdf, 0) (the last bin of the reference's spectra is negative)
    std::vector<float> cdf(c->n_bins);
    double tot = 0.0;
    for (int b = 0; b < c->n_bins; b++) tot += c->h_pdf[b] > 0 ? c->h_pdf[b] : 0.0;
    double run = 0.0;
    for (int b = 0; b < c->n_bins; b++) { run += c->h_pdf[b] > 0 ? c->h_pdf[b] : 0.0; cdf[b] = (float)(run / tot); }
    cdf[c->n_bins - 1] = 1.0f;
    cudaFree(c->d_sc_cdf); c->d_sc_cdf = nullptr;
    CU(c, cudaMalloc(&c->d_sc_cdf, sizeof(float) * c->n_bins));
    CU(c, cudaMemcpy(c->d_sc_cdf, cdf.data(), sizeof(float) * c->n_bins, cudaMemcpyHostToDevice));
    if (!c->d_sc_counters) CU(c, cudaMalloc(&c->d_sc_counters, sizeof(double) * 8));
    c->sc_ready = true;
    return DRR_OK;
}

int drr_scatter(drr_ctx* c, unsigned long long n_photons, unsigned long long photon_offset, uint64_t seed, int W, int H, const float* w2i,
                const float* index_from_world, const float* source_world, const float* ijk_from_world, unsigned long long* out_tally,
                double* out_counters, int out_mem_kind) {
    if (!c) return DRR_E_INVALID;
    if (!c->sc_ready) return fail(c, DRR_E_STATE, "drr_scatter: call drr_set_scatter_tables first");
    const int V = (int)c->vols.size();
    if (V < 1) return fail(c, DRR_E_STATE, "drr_scatter: add a volume first");
    for (int v = 0; v < V; v++)
        if (!c->vols[v].dens) return fail(c, DRR_E_STATE, "drr_scatter: volume %d has no raw arrays", v);
    if (W <= 0 || H <= 0 || !w2i || !index_from_world || !source_world || !ijk_from_world || !out_tally)
        return fail(c, DRR_E_INVALID, "drr_scatter: bad arguments");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t npix = (size_t)W * H;
    int rc;
    if ((rc = ensure(c, (void**)&c->d_sc_tally, &c->sc_tally_cap, sizeof(unsigned long long) * npix))) return rc;
    CU(c, cudaMemsetAsync(c->d_sc_tally, 0, sizeof(unsigned long long) * npix, s));
    CU(c, cudaMemsetAsync(c->d_sc_counters, 0, sizeof(double) * 8, s));
    ScatterParams P;
    memset(&P, 0, sizeof P);
    P.T = c->sc;
    P.V = V;
    for (int v = 0; v < V; v++) {
        const VolHost& h = c->vols[v];
        P.vol[v].dens = h.dens; P.vol[v].lab = h.lab; P.vol[v].ni = h.ni; P.vol[v].nj = h.nj; P.vol[v].nk = h.nk;
        P.priority[v] = c->priorities_set ? c->priority[v] : V - 1 - v;  // projector.py:489-492
        P.enabled[v] = c->enabled[v];
        memcpy(P.ijk[v], ijk_from_world + 12 * v, 48);
    }
    memcpy(P.p_idx, index_from_world, 48); memcpy(P.w2i, w2i, 36); memcpy(P.src, source_world, 12);
    P.W = W; P.H = H; P.n_bins = c->n_bins; P.spec_e_keV = c->d_energies; P.spec_cdf = c->d_sc_cdf;
    P.n_photons = n_photons; P.photon_offset = photon_offset; P.seed = seed;
    P.tally = c->d_sc_tally; P.counters = c->d_sc_counters;
    CU(c, cudaEventRecord(c->ev[1], s));
    if (n_photons > 0) { CU(c, drr_launch_scatter(P, c->n_sm, s)); c->launches += 1; }
    CU(c, cudaEventRecord(c->ev[2], s));
    const cudaMemcpyKind kind = out_mem_kind == DRR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    CU(c, cudaMemcpyAsync(out_tally, c->d_sc_tally, sizeof(unsigned long long) * npix, kind, s));
    if (out_counters) CU(c, cudaMemcpyAsync(out_counters, c->d_sc_counters, sizeof(double) * 8, cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    CU(c, cudaEventElapsedTime(&c->last_ms[0], c->ev[1], c->ev[2]));
    return DRR_OK;
}

int drr_mesh_query(drr_ctx* c, int mode, int W, int H, const float* w2i, const uint8_t* select, void* out, int mem_kind) {
    if (!c) return DRR_E_INVALID;
    if (c->n_prims == 0) return fail(c, DRR_E_STATE, "drr_mesh_query: call drr_set_meshes first");
    if (c->pose_views < 1) return fail(c, DRR_E_STATE, "drr_mesh_query: call drr_set_mesh_poses first");
    if (mode < 0 || mode > 2 || W <= 0 || H <= 0 || !w2i || !select || !out) return fail(c, DRR_E_INVALID, "drr_mesh_query: bad arguments");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t npix = (size_t)W * H;
    const int MH = c->own_max_hits, np = c->n_prims;
    // a private copy of the primitive table with the selection folded into the flags the tracing kernels test
    std::vector<MeshPrimDev> prims(c->h_prims);
    for (int p = 0; p < np; p++) {
        MeshPrimDev& d = prims[p];
        const bool sel = select[p] != 0;
        if (mode == DRR_MESH_QUERY_TRAVEL) {
            d.additive = (sel && d.additive && d.layer == 0) ? 1 : 0;  // renderer.py:312-319 with layer_idx=0
            d.subtractive = 0; d.mat_slot = 0; d.density = 1.0f; d.layer = 0;
        } else {
            d.additive = 0; d.subtractive = sel ? 1 : 0; d.layer = 0;
        }
    }
    const size_t out_bytes = mode == DRR_MESH_QUERY_HITS ? npix * MH * 4 : (mode == DRR_MESH_QUERY_TRAVEL ? npix * 4 : npix);
    const size_t scratch_bytes = mode == DRR_MESH_QUERY_HITS ? npix * MH : (mode == DRR_MESH_QUERY_TRAVEL ? npix * 8 : 0);
    MeshPrimDev* d_prims = nullptr; ViewDev* d_view = nullptr; float *d_wfm = nullptr, *d_src = nullptr, *d_vw = nullptr;
    void *d_out = nullptr, *d_scratch = nullptr; int8_t* d_valid = nullptr; int4* d_tbox = nullptr; int* d_pbox = nullptr;
    int rc = DRR_OK;
    ViewDev hv; memset(&hv, 0, sizeof hv); memcpy(hv.w2i, w2i, 36); invert3(hv.w2i, hv.w2i_inv);
#define QCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = fail(c, DRR_E_CUDA, "drr_mesh_query: %s", cudaGetErrorString(e_)); goto done; } } while (0)
    QCU(cudaMalloc(&d_prims, sizeof(MeshPrimDev) * np));
    QCU(cudaMalloc(&d_view, sizeof(ViewDev)));
    QCU(cudaMalloc(&d_wfm, sizeof(float) * 12 * np));
    QCU(cudaMalloc(&d_src, 12));
    QCU(cudaMalloc(&d_vw, sizeof(float) * 9 * (size_t)c->n_tris));
    QCU(cudaMalloc(&d_valid, 1));
    QCU(cudaMalloc(&d_tbox, sizeof(int4) * (size_t)(c->n_tris > 0 ? c->n_tris : 1)));
    QCU(cudaMalloc(&d_pbox, sizeof(int) * 4 * np));
    if (scratch_bytes) QCU(cudaMalloc(&d_scratch, scratch_bytes));
    if (mem_kind == DRR_MEM_HOST) QCU(cudaMalloc(&d_out, out_bytes)); else d_out = out;
    QCU(cudaMemcpyAsync(d_prims, prims.data(), sizeof(MeshPrimDev) * np, cudaMemcpyHostToDevice, s));
    QCU(cudaMemcpyAsync(d_view, &hv, sizeof hv, cudaMemcpyHostToDevice, s));
    QCU(cudaMemcpyAsync(d_wfm, c->h_world_from_mesh.data(), sizeof(float) * 12 * np, cudaMemcpyHostToDevice, s));
    QCU(cudaMemcpyAsync(d_src, c->h_source_world.data(), 12, cudaMemcpyHostToDevice, s));
    QCU(cudaMemsetAsync(d_valid, 0, 1, s));
    QCU(drr_launch_mesh_transform(c->d_verts_local, c->d_prim_of_tri, d_wfm, c->n_tris, np, 1, d_vw, s));
    QCU(drr_launch_mesh_project(d_view, d_src, d_vw, c->d_prim_of_tri, c->n_tris, np, 1, d_tbox, d_pbox, s));
    c->launches += 3;
    if (mode == DRR_MESH_QUERY_HITS) {
        QCU(drr_launch_mesh_subtractive(d_view, d_src, d_vw, d_prims, np, c->n_tris, 0, 1, W, H, 1, MH, c->far_limit, (float*)d_out,
                                        (int8_t*)d_scratch, d_tbox, d_pbox, s));
        c->launches += 1;
    } else if (mode == DRR_MESH_QUERY_TRAVEL) {
        QCU(cudaMemsetAsync(d_scratch, 0, scratch_bytes, s));
        QCU(drr_launch_mesh_additive(d_view, d_src, d_vw, d_prims, np, c->n_tris, 1, 1, W, H, 1, MH, d_valid, nullptr, nullptr, (float*)d_scratch,
                                     d_tbox, d_pbox, s));
        QCU(drr_launch_mesh_travel_finish((const float*)d_scratch, (int)npix, (float*)d_out, s));
        c->launches += 2;
    } else {
        QCU(drr_launch_mesh_cover(d_view, d_src, d_vw, d_prims, np, W, H, (uint8_t*)d_out, d_tbox, d_pbox, s));
        c->launches += 1;
    }
    if (mem_kind == DRR_MEM_HOST) QCU(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    QCU(cudaStreamSynchronize(s));
#undef QCU
done:
    cudaFree(d_prims); cudaFree(d_view); cudaFree(d_wfm); cudaFree(d_src); cudaFree(d_vw); cudaFree(d_valid); cudaFree(d_scratch);
    cudaFree(d_tbox); cudaFree(d_pbox);
    if (mem_kind == DRR_MEM_HOST) cudaFree(d_out);
    return rc;
}

int drr_mesh_clean_hits(drr_ctx* c, float* ts, int8_t* facing, int n_rays, int n, float far_limit, int mem_kind) {
    if (!c) return DRR_E_INVALID;
    if (!ts || !facing || n_rays <= 0 || n <= 0 || n > 128) return fail(c, DRR_E_INVALID, "drr_mesh_clean_hits: bad arguments");
    CU(c, cudaSetDevice(c->device));
    float* dt = ts; int8_t* df = facing;
    const size_t cnt = (size_t)n_rays * n;
    if (mem_kind == DRR_MEM_HOST) {
        CU(c, cudaMalloc(&dt, cnt * 4)); CU(c, cudaMalloc(&df, cnt));
        CU(c, cudaMemcpyAsync(dt, ts, cnt * 4, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(df, facing, cnt, cudaMemcpyHostToDevice, c->stream));
    }
    CU(c, drr_launch_tide_clean(dt, df, n_rays, n, far_limit, c->stream));
    c->launches += 1;
    if (mem_kind == DRR_MEM_HOST) {
        CU(c, cudaMemcpyAsync(ts, dt, cnt * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(facing, df, cnt, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    if (mem_kind == DRR_MEM_HOST) { cudaFree(dt); cudaFree(df); }
    return DRR_OK;
}

// The warp-cooperative kernel stages the voxel cells an 8x4-pixel tile touches; it pays when neighbouring
// rays are closer than a few voxels.  Estimate the tile's footprint at the volume centre for view 0 and
// fall back to the per-ray kernel for coarse detectors / strongly magnified set-ups.
#ifndef PER_RAY_MIN_SPREAD
#define PER_RAY_MIN_SPREAD 4.0f  // voxels across an 8 x 4 pixel tile beyond which the per-ray kernel takes over
#endif
#define PER_RAY_MIN_STEP_VOX 0.9f  // voxels per step beyond which the per-ray kernel takes over
#define RAYS2_MAX_SPREAD 2.2f  // voxels across an 8 x 4 pixel tile up to which the single-volume march walks two rays per lane

static float tile_spread(const drr_ctx* c, const float* w2i, const float* src, const float* ijk, int W, int H, float* vox_per_mm = nullptr) {
    const VolHost& v = c->vols[0];
    auto dir = [&](float u, float vv, float* d) {
        float r[3];
        for (int a = 0; a < 3; a++) r[a] = u * w2i[3 * a] + vv * w2i[3 * a + 1] + w2i[3 * a + 2];
        float len = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        for (int a = 0; a < 3; a++) d[a] = (ijk[4 * a] * r[0] + ijk[4 * a + 1] * r[1] + ijk[4 * a + 2] * r[2]) / len;
    };
    float d0[3], d1[3];
    dir(0.5f * W, 0.5f * H, d0);
    dir(0.5f * W + 8.0f, 0.5f * H + 4.0f, d1);
    const float ctr[3] = {0.5f * v.ni - src[0], 0.5f * v.nj - src[1], 0.5f * v.nk - src[2]};
    float dd = d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2];
    float alpha_c = dd > 0 ? (ctr[0] * d0[0] + ctr[1] * d0[1] + ctr[2] * d0[2]) / dd : 0.0f;
    float spread = 0.0f;
    for (int a = 0; a < 3; a++) spread = fmaxf(spread, fabsf(alpha_c * (d1[a] - d0[a])));
    if (vox_per_mm) *vox_per_mm = sqrtf(dd);  // voxels per mm along the central ray
    return spread;
}

static int pick_variant(const drr_ctx* c, const float* w2i, const float* src, const float* ijk, int W, int H) {
    if (c->variant != 0) return c->variant;
    float vox_per_mm = 0.0f;
    const float spread = tile_spread(c, w2i, src, ijk, W, H, &vox_per_mm);
    // more than ~4 voxels across a tile, or steps of about a voxel and more (the 32 steps of a segment then cover so many cells that
    // the staged boxes have to be cut down again and again: C2 at a step of 1 mm, 1.25 voxels, 5.5 against 4.8 ms per view; at 2 mm
    // 5.7 against 2.5; at 0.5 mm the lock-step kernel is still 1.7 x faster -- tools/step_sweep.py): per-ray kernel
    return (spread > PER_RAY_MIN_SPREAD || c->step * vox_per_mm > PER_RAY_MIN_STEP_VOX) ? 1 : 0;
}

// Slack of the lock-step kernels' staged boxes and window tests for this batch (see drr_march_warp.cu).  A segment of S steps is
// bounded from alpha and fma(S, step, alpha); the alpha the kernel really reaches after j <= S sequential fp32 additions can be
// (S / 2 + 1) ulps of the largest alpha away from that, i.e. that many ulps times |row of ijk_from_world| voxels, plus the
// roundings of the coordinate FMAs themselves.  C2 (alpha <= 1100 mm, 0.8 mm voxels) needs 0.003 voxel; a 0.1 mm K-wire grid seen
// from a metre away 0.05.  Returns false when a volume would need more than a quarter voxel: such scenes take the per-ray kernels.
static bool march_slack(const drr_ctx* c, int n_views, const float* src_ijk, const float* ijk_from_world, float max_ray_length,
                        MarchParams& P) {
    const int V = (int)c->vols.size();
    const double S = 64.0;  // the longest segment of drr_march_warp.cu (SEG_TEX)
    double worst_vox = 0.0, worst_a = 0.0;
    for (int i = 0; i < n_views; i++)
        for (int v = 0; v < V; v++) {
            const float* A = ijk_from_world + ((size_t)i * V + v) * 12;
            const float* sp = src_ijk + ((size_t)i * V + v) * 3;
            const VolHost& h = c->vols[v];
            const float a9[9] = {A[0], A[1], A[2], A[4], A[5], A[6], A[8], A[9], A[10]};
            float inv[9];
            invert3(a9, inv);
            double alpha_max = 0.0;
            const double hi[3] = {h.ni - 0.5, h.nj - 0.5, h.nk - 0.5};
            for (int k = 0; k < 8; k++) {
                const double d[3] = {((k & 1) ? hi[0] : -0.5) - sp[0], ((k & 2) ? hi[1] : -0.5) - sp[1], ((k & 4) ? hi[2] : -0.5) - sp[2]};
                double w2 = 0.0;
                for (int r = 0; r < 3; r++) {
                    const double w = inv[3 * r] * d[0] + inv[3 * r + 1] * d[1] + inv[3 * r + 2] * d[2];
                    w2 += w * w;
                }
                alpha_max = fmax(alpha_max, sqrt(w2));
            }
            if (!(alpha_max > 0.0) || !std::isfinite(alpha_max)) alpha_max = 1.0e9;  // singular pose matrix: no bound
            if (max_ray_length > 0.0f) alpha_max = fmin(alpha_max, (double)max_ray_length);
            alpha_max = fmax(alpha_max, 2.0);  // rays start at alpha ~ 1 (ray_length)
            const double ulp_a = ldexp(1.0, (int)floor(log2(alpha_max)) - 23);
            double row = 0.0;
            for (int r = 0; r < 3; r++) row = fmax(row, sqrt((double)a9[3 * r] * a9[3 * r] + (double)a9[3 * r + 1] * a9[3 * r + 1] + (double)a9[3 * r + 2] * a9[3 * r + 2]));
            const double nmax = fmax(fmax(h.ni, h.nj), h.nk) + 2.0;
            const double ulp_p = ldexp(1.0, (int)floor(log2(nmax)) - 23);
            worst_vox = fmax(worst_vox, (S / 2 + 1) * ulp_a * row + 4.0 * ulp_p);
            worst_a = fmax(worst_a, (S / 2 + 1) * ulp_a);
        }
    const double lo = fmax(0.01, 0.002 + 1.25 * worst_vox);
    P.slack_lo = (float)lo;
    P.slack_hi = (float)(lo + 0.0025);          // + the 1 / 512 round-up of the fixed-point coordinate
    P.slack_alpha = (float)fmax(0.01, 1.5 * worst_a);
    return lo <= 0.25 && P.slack_alpha <= 0.25f * c->step + 0.01f;
}

static int ensure(drr_ctx* c, void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *p) return DRR_OK;
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    CU(c, cudaMalloc(p, bytes));
    *cap = bytes;
    return DRR_OK;
}

int drr_project(drr_ctx* c, int n_views, int W, int H, const float* w2i, const float* src_ijk, const float* ijk_from_world,
                float max_ray_length, unsigned post_flags, float photon_count, float intensity_upper_bound, float pixel_area_mm2,
                uint64_t seed, float* out_intensity, float* out_pprob, float* out_area, int out_mem_kind) {
    if (!c) return DRR_E_INVALID;
    if (n_views <= 0 || W <= 0 || H <= 0 || !w2i) return fail(c, DRR_E_INVALID, "drr_project: bad view / sensor arguments");
    if (c->n_bins == 0) return fail(c, DRR_E_STATE, "drr_project: call drr_set_spectrum first");
    const int V = (int)c->vols.size();
    if (V > 0 && (!src_ijk || !ijk_from_world)) return fail(c, DRR_E_INVALID, "drr_project: missing per-volume pose arrays");
    if (!out_intensity && !out_area) return fail(c, DRR_E_INVALID, "drr_project: no output requested");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t npix = (size_t)W * H;
    const int M = c->M;

    // per-view poses -> pinned staging -> device (one copy per batch instead of 5 per view, projector.py:802-831)
    if (c->views_cap < n_views) {
        cudaFree(c->d_views); cudaFreeHost(c->h_views);
        c->d_views = nullptr; c->h_views = nullptr; c->views_cap = 0;
        CU(c, cudaMalloc(&c->d_views, sizeof(ViewDev) * n_views));
        CU(c, cudaMallocHost(&c->h_views, sizeof(ViewDev) * n_views));
        c->views_cap = n_views;
    }
    for (int i = 0; i < n_views; i++) {
        ViewDev& vd = c->h_views[i];
        memset(&vd, 0, sizeof vd);
        memcpy(vd.w2i, w2i + (size_t)i * 9, 36);
        invert3(vd.w2i, vd.w2i_inv);
        for (int v = 0; v < V; v++) {
            memcpy(vd.src[v], src_ijk + ((size_t)i * V + v) * 3, 12);
            memcpy(vd.ijk[v], ijk_from_world + ((size_t)i * V + v) * 12, 48);
        }
    }
    CU(c, cudaEventRecord(c->ev[0], s));
    CU(c, cudaMemcpyAsync(c->d_views, c->h_views, sizeof(ViewDev) * n_views, cudaMemcpyHostToDevice, s));

    int rc;
    if ((rc = ensure(c, (void**)&c->d_area, &c->area_cap, sizeof(float) * npix * M * n_views))) return rc;
    if ((rc = ensure(c, (void**)&c->d_intensity, &c->int_cap, sizeof(float) * npix * n_views))) return rc;
    if ((rc = ensure(c, (void**)&c->d_pprob, &c->pp_cap, sizeof(float) * npix * n_views))) return rc;
    CU(c, cudaMemsetAsync(c->d_samples, 0, 2 * sizeof(unsigned long long), s));
    CU(c, cudaMemsetAsync(c->d_tile_counter, 0, sizeof(unsigned int), s));

    MarchParams P;
    memset(&P, 0, sizeof P);
    for (int v = 0; v < V; v++) {
        const VolHost& h = c->vols[v];
        P.vol[v].dens = h.dens; P.vol[v].lab = h.lab; P.vol[v].cellc = h.cellc; P.vol[v].celll = h.celll; P.vol[v].cellcode = h.cellcode;
        P.vol[v].tex = h.tex; P.vol[v].ni = h.ni; P.vol[v].nj = h.nj; P.vol[v].nk = h.nk;
        P.priority[v] = c->priorities_set ? c->priority[v] : V - 1 - v;  // projector.py:489-492
        P.enabled[v] = c->enabled[v];
    }
    P.V = V; P.M = M; P.W = W; P.H = H; P.n_views = n_views;
    P.step = c->step; P.max_ray_length = max_ray_length;
    P.attenuate_outside = c->attenuate_outside; P.air_index = c->air_index;
    P.mesh_layers = c->mesh_layers; P.max_hits = c->max_hits; P.n_mesh_mats = c->n_mesh_mats;
    P.hit_alphas = c->hit_alphas; P.hit_facing = c->hit_facing; P.layer_valid = c->layer_valid;
    P.additive = c->additive; P.mesh_mats = c->mesh_mats;
    P.views = c->d_views; P.area = c->d_area; P.sample_count = c->d_samples; P.tile_counter = c->d_tile_counter;

    // ---- meshes traced by this library: fill hit lists + additive buffers for this batch --------------
    if (c->n_prims > 0) {
        if (c->pose_views != n_views) return fail(c, DRR_E_STATE, "drr_project: call drr_set_mesh_poses for this batch of %d views first", n_views);
        const int L = c->own_layers, MH = c->own_max_hits, NMm = c->own_n_mats;
        if ((rc = ensure(c, (void**)&c->d_world_from_mesh, &c->wfm_cap, sizeof(float) * 12 * (size_t)n_views * c->n_prims))) return rc;
        if ((rc = ensure(c, (void**)&c->d_source_world, &c->srcw_cap, sizeof(float) * 3 * (size_t)n_views))) return rc;
        if ((rc = ensure(c, (void**)&c->d_verts_world, &c->vw_cap, sizeof(float) * 9 * (size_t)n_views * c->n_tris))) return rc;
        if ((rc = ensure(c, (void**)&c->d_own_hit_alphas, &c->oha_cap, sizeof(float) * (size_t)n_views * L * npix * MH))) return rc;
        if ((rc = ensure(c, (void**)&c->d_own_hit_facing, &c->ohf_cap, (size_t)n_views * L * npix * MH))) return rc;
        if ((rc = ensure(c, (void**)&c->d_own_additive, &c->oadd_cap, sizeof(float) * 2 * (size_t)n_views * L * NMm * npix))) return rc;
        CU(c, cudaMemcpyAsync(c->d_world_from_mesh, c->h_world_from_mesh.data(), sizeof(float) * 12 * (size_t)n_views * c->n_prims, cudaMemcpyHostToDevice, s));
        CU(c, cudaMemcpyAsync(c->d_source_world, c->h_source_world.data(), sizeof(float) * 3 * (size_t)n_views, cudaMemcpyHostToDevice, s));
        CU(c, cudaMemsetAsync(c->d_own_hit_facing, 0, (size_t)n_views * L * npix * MH, s));
        CU(c, cudaMemsetAsync(c->d_own_additive, 0, sizeof(float) * 2 * (size_t)n_views * L * NMm * npix, s));
        if ((rc = ensure(c, (void**)&c->d_tri_box, &c->tribox_cap, sizeof(int4) * (size_t)n_views * (c->n_tris > 0 ? c->n_tris : 1)))) return rc;
        if ((rc = ensure(c, (void**)&c->d_prim_box, &c->primbox_cap, sizeof(int) * 4 * (size_t)n_views * c->n_prims))) return rc;
        CU(c, drr_launch_mesh_transform(c->d_verts_local, c->d_prim_of_tri, c->d_world_from_mesh, c->n_tris, c->n_prims, n_views, c->d_verts_world, s));
        CU(c, drr_launch_mesh_project(c->d_views, c->d_source_world, c->d_verts_world, c->d_prim_of_tri, c->n_tris, c->n_prims, n_views,
                                      c->d_tri_box, c->d_prim_box, s));
        c->launches += 3;
        for (int l = L - 1; l >= 0; l--) {
            if (!c->h_layer_valid[l]) continue;
            CU(c, drr_launch_mesh_subtractive(c->d_views, c->d_source_world, c->d_verts_world, c->d_prims, c->n_prims, c->n_tris, l, L, W, H, n_views,
                                              MH, c->far_limit, c->d_own_hit_alphas, c->d_own_hit_facing, c->d_tri_box, c->d_prim_box, s));
            c->launches += 1;
        }
        CU(c, drr_launch_mesh_additive(c->d_views, c->d_source_world, c->d_verts_world, c->d_prims, c->n_prims, c->n_tris, L, NMm, W, H, n_views, MH,
                                       c->d_own_layer_valid, c->d_own_hit_alphas, c->d_own_hit_facing, c->d_own_additive, c->d_tri_box,
                                       c->d_prim_box, s));
        c->launches += 1;
        P.mesh_layers = L; P.max_hits = MH; P.n_mesh_mats = NMm;
        P.hit_alphas = c->d_own_hit_alphas; P.hit_facing = c->d_own_hit_facing; P.layer_valid = c->d_own_layer_valid;
        P.additive = c->d_own_additive; P.mesh_mats = c->d_own_mesh_mats;
    }
    const bool meshes = P.layer_valid || P.additive;
    bool single = (V == 1) && !meshes && !c->attenuate_outside && M <= 8;
    if (single) {
        const VolHost& h = c->vols[0];
        int sampler = c->sampler;
        if (sampler != DRR_SAMPLER_TEX && !h.cellc) sampler = h.tex ? DRR_SAMPLER_TEX : -1;
        if (sampler != DRR_SAMPLER_ALU && !h.tex) sampler = h.cellc ? DRR_SAMPLER_ALU : -1;
        if (sampler < 0) return fail(c, DRR_E_STATE, "drr_project: volume 0 has neither cell records nor a texture");
        P.tex_eighths = sampler == DRR_SAMPLER_ALU ? 0 : (sampler == DRR_SAMPLER_TEX ? 8 : c->tex_eighths);
    }
    const bool lockstep_ok = V > 0 && march_slack(c, n_views, src_ijk, ijk_from_world, max_ray_length, P);
    P.lane_quads = c->lane_quads == 2 ? 1 : c->lane_quads;  // 2 x 2 lane groups (see drr_march_warp.cu: lane_u); single-volume kernel only
    // two rays per lane while an 8 x 4 pixel tile spans at most ~2 voxels (C2 at 1536^2: 1.0; crossover between 768^2 and 640^2)
    P.rays_per_lane = c->rays_per_lane ? c->rays_per_lane : ((single && tile_spread(c, w2i, src_ijk, ijk_from_world, W, H) <= RAYS2_MAX_SPREAD) ? 2 : 1);
    // ---- host-bound single-volume batches: two halves, the device-to-host copy of the first under the march of the second -----
    // (projector.py:786-792 copies every view back before the next one starts; here only the second half's copy is exposed)
    const bool lockstep_single = single && h_has_cells(c) && (c->variant == 0 ? lockstep_ok : true) &&
                                 pick_variant(c, w2i, src_ijk, ijk_from_world, W, H) == 0;
    if (lockstep_single && c->pipeline && out_intensity && !out_area && !out_pprob && out_mem_kind == DRR_MEM_HOST && n_views >= 4 &&
        !(post_flags & (DRR_POST_NOISE | DRR_POST_COLLECTED))) {
        if (c->minmax_cap < n_views) {
            cudaFree(c->d_minmax); cudaFree(c->d_viewsum);
            CU(c, cudaMalloc(&c->d_minmax, sizeof(unsigned) * 2 * n_views));
            CU(c, cudaMalloc(&c->d_viewsum, sizeof(double) * n_views));
            c->minmax_cap = n_views;
        }
        CU(c, cudaMemsetAsync(c->d_const_flag, 0, sizeof(int), s));
        const int first = n_views - (c->pipeline < n_views / 2 ? c->pipeline : n_views / 2);  // the last piece's copy is the exposed one
        for (int k = 0; k < 2; k++) {
            const int v0 = k ? first : 0, nv = k ? n_views - first : first;
            MarchParams Pk = P;
            Pk.views = c->d_views + v0; Pk.area = c->d_area + (size_t)v0 * M * npix; Pk.n_views = nv;
            if (k) CU(c, cudaMemsetAsync(c->d_tile_counter, 0, sizeof(unsigned int), s));
            CU(c, cudaEventRecord(c->evp[3 * k], s));
            CU(c, drr_launch_march_warp(Pk, c->n_sm, s));
            CU(c, cudaEventRecord(c->evp[3 * k + 1], s));
            float* img = c->d_intensity + (size_t)v0 * npix;
            CU(c, drr_launch_spectral(Pk.area, c->n_bins, M, c->d_energies, c->d_pdf, c->d_mu, npix, nv, img, c->d_pprob + (size_t)v0 * npix, c->n_sm, s));
            c->launches += 2;
            if (post_flags & DRR_POST_CLIP) { CU(c, drr_launch_clip(img, npix * nv, intensity_upper_bound, s)); c->launches += 1; }
            if (post_flags & DRR_POST_NEGLOG) { CU(c, drr_launch_neglog(img, npix, nv, c->d_minmax + 2 * v0, 0.01f, s, c->d_const_flag)); c->launches += 2; }
            CU(c, cudaEventRecord(c->evp[3 * k + 2], s));
        }
        // the copies are queued after ALL the compute work: a pageable destination makes cudaMemcpyAsync block the host until its
        // copy is done, and the second piece must already be on the GPU's queue by then
        for (int k = 0; k < 2; k++) {
            const int v0 = k ? first : 0, nv = k ? n_views - first : first;
            CU(c, cudaStreamWaitEvent(c->copy_stream, c->evp[3 * k + 2], 0));
            CU(c, cudaMemcpyAsync(out_intensity + (size_t)v0 * npix, c->d_intensity + (size_t)v0 * npix, sizeof(float) * npix * nv,
                                  cudaMemcpyDeviceToHost, c->copy_stream));
        }
        CU(c, cudaEventRecord(c->evp[6], c->copy_stream));
        CU(c, cudaMemcpyAsync(c->last_samples, c->d_samples, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CU(c, cudaMemcpyAsync(&c->h_const_flag, c->d_const_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(c, cudaStreamSynchronize(s));
        CU(c, cudaStreamSynchronize(c->copy_stream));
        // utils.neglog zeroes the WHOLE batch when any image of it is constant (image_utils.py:42-49): a piece that met one has
        // zeroed itself; tell the other piece
        if ((post_flags & DRR_POST_NEGLOG) && c->h_const_flag) memset(out_intensity, 0, sizeof(float) * npix * n_views);
        float m0 = 0, m1 = 0, p0 = 0, p1 = 0, tot = 0;
        CU(c, cudaEventElapsedTime(&m0, c->evp[0], c->evp[1])); CU(c, cudaEventElapsedTime(&m1, c->evp[3], c->evp[4]));
        CU(c, cudaEventElapsedTime(&p0, c->evp[1], c->evp[2])); CU(c, cudaEventElapsedTime(&p1, c->evp[4], c->evp[5]));
        CU(c, cudaEventElapsedTime(&tot, c->ev[0], c->evp[6]));
        c->last_ms[0] = m0 + m1; c->last_ms[1] = p0 + p1; c->last_ms[2] = tot;
        return DRR_OK;
    }
    CU(c, cudaEventRecord(c->ev[1], s));
    if (V == 0) {
        if (meshes) { CU(c, drr_launch_march_meshonly(P, s)); c->launches += 1; }
        else CU(c, cudaMemsetAsync(c->d_area, 0, sizeof(float) * npix * M * n_views, s));
    } else if (single) {
        if (lockstep_single) {
            CU(c, drr_launch_march_warp(P, c->n_sm, s));  // persistent: every warp pulls 8x4-pixel tiles from the queue
        } else {
            int occ = drr_march_single_occupancy(M);
            if (occ < 1) occ = 1;
            CU(c, drr_launch_march_single(P, c->n_sm * occ, s));
        }
        c->launches += 1;
    } else {
        if (V > DRR_MAX_VOLUMES || M > DRR_MAX_MATERIALS)
            return fail(c, DRR_E_INVALID, "drr_project: at most %d volumes / %d materials", DRR_MAX_VOLUMES, DRR_MAX_MATERIALS);
        for (int v = 0; v < V; v++)
            if (!c->vols[v].dens) return fail(c, DRR_E_STATE, "drr_project: volume %d has no raw arrays", v);
        // Tiles whose rays see a single volume take the lock-step kernel; the others are listed for the general one.
        // (V == 1 gets here only with meshes or more than 8 materials; the lock-step kernels are built for V <= 4, M <= 8)
        bool split = !c->attenuate_outside && c->variant == 0 && lockstep_ok && V <= 4 && M <= 8;
        int sampler = c->sampler;
        for (int v = 0; v < V; v++)  // a volume without coefficient records: texture unit only; without a texture: FMA pipes only
            if (!c->vols[v].cellc && c->vols[v].tex) sampler = DRR_SAMPLER_TEX;
        for (int v = 0; v < V; v++)
            if (!c->vols[v].tex && sampler != DRR_SAMPLER_TEX) sampler = DRR_SAMPLER_ALU;
        for (int v = 0; v < V && split; v++) {
            const VolHost& h = c->vols[v];
            if (!h.cellcode || (sampler != DRR_SAMPLER_TEX && !h.cellc) || (sampler != DRR_SAMPLER_ALU && !h.tex)) split = false;
            // Volumes that share a priority all contribute whenever one of them is picked, even outside their own
            // box (K.cu:540-547 tests the priority only): such scenes are replayed step by step.
            for (int u = 0; u < v; u++)
                if (P.enabled[u] && P.enabled[v] && P.priority[u] == P.priority[v]) split = false;
        }
        if (split) {
            const size_t n_tiles = (size_t)((W + 7) / 8) * ((H + 3) / 4) * n_views;
            if ((rc = ensure(c, (void**)&c->d_worklist, &c->worklist_cap, sizeof(unsigned) * (n_tiles + 2)))) return rc;
            CU(c, cudaMemsetAsync(c->d_worklist, 0, sizeof(unsigned) * 2, s));
            P.work_count = c->d_worklist; P.worklist = c->d_worklist + 2;
            P.tex_eighths = sampler == DRR_SAMPLER_ALU ? 0 : (sampler == DRR_SAMPLER_TEX ? 8 : c->tex_eighths);
            CU(c, drr_launch_march_multi(P, c->n_sm, s));
            CU(c, drr_launch_march_general_list(P, c->n_sm * 8, s));
            c->launches += 2;
        } else {
            CU(c, drr_launch_march_general(P, s));
            c->launches += 1;
        }
    }
    CU(c, cudaEventRecord(c->ev[2], s));

    if (out_intensity) {
        CU(c, drr_launch_spectral(c->d_area, c->n_bins, M, c->d_energies, c->d_pdf, c->d_mu, npix, n_views, c->d_intensity, c->d_pprob,
                                  c->n_sm, s));
        c->launches += 1;
        if (post_flags & DRR_POST_COLLECTED) {
            if ((rc = ensure(c, (void**)&c->d_scratch, &c->scratch_cap, sizeof(float) * npix * n_views))) return rc;
            if (c->minmax_cap < n_views) {
                cudaFree(c->d_minmax); cudaFree(c->d_viewsum);
                CU(c, cudaMalloc(&c->d_minmax, sizeof(unsigned) * 2 * n_views));
                CU(c, cudaMalloc(&c->d_viewsum, sizeof(double) * n_views));
                c->minmax_cap = n_views;
            }
            CU(c, drr_launch_collected(c->d_intensity, c->d_scratch, c->d_viewsum, c->d_views, W, H, n_views, photon_count, pixel_area_mm2, s));
            c->launches += 2;
        }
        if (post_flags & DRR_POST_NOISE) {
            if ((rc = ensure(c, (void**)&c->d_scratch, &c->scratch_cap, sizeof(float) * npix * n_views))) return rc;
            CU(c, drr_launch_noise(c->d_intensity, c->d_pprob, c->d_scratch, W, H, n_views, photon_count, seed, s));
            c->launches += 2;
        }
        if (post_flags & DRR_POST_CLIP) {
            CU(c, drr_launch_clip(c->d_intensity, npix * n_views, intensity_upper_bound, s));
            c->launches += 1;
        }
        if (post_flags & DRR_POST_NEGLOG) {
            if (c->minmax_cap < n_views) {
                cudaFree(c->d_minmax); cudaFree(c->d_viewsum);
                CU(c, cudaMalloc(&c->d_minmax, sizeof(unsigned) * 2 * n_views));
                CU(c, cudaMalloc(&c->d_viewsum, sizeof(double) * n_views));
                c->minmax_cap = n_views;
            }
            CU(c, drr_launch_neglog(c->d_intensity, npix, n_views, c->d_minmax, 0.01f, s));
            c->launches += 2;
        }
    }
    CU(c, cudaEventRecord(c->ev[3], s));
    const cudaMemcpyKind kind = out_mem_kind == DRR_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (out_intensity) CU(c, cudaMemcpyAsync(out_intensity, c->d_intensity, sizeof(float) * npix * n_views, kind, s));
    if (out_pprob && out_intensity) CU(c, cudaMemcpyAsync(out_pprob, c->d_pprob, sizeof(float) * npix * n_views, kind, s));
    if (out_area) CU(c, cudaMemcpyAsync(out_area, c->d_area, sizeof(float) * npix * M * n_views, kind, s));
    CU(c, cudaMemcpyAsync(c->last_samples, c->d_samples, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(c, cudaEventRecord(c->ev[4], s));
    CU(c, cudaStreamSynchronize(s));
    CU(c, cudaEventElapsedTime(&c->last_ms[0], c->ev[1], c->ev[2]));
    CU(c, cudaEventElapsedTime(&c->last_ms[1], c->ev[2], c->ev[3]));
    CU(c, cudaEventElapsedTime(&c->last_ms[2], c->ev[0], c->ev[4]));
    return DRR_OK;
}

int drr_postprocess(drr_ctx* c, float* images, const float* photon_prob, int n_views, int W, int H, unsigned post_flags, float photon_count,
                    float intensity_upper_bound, uint64_t seed, int mem_kind) {
    if (!c) return DRR_E_INVALID;
    if (!images || n_views <= 0 || W <= 0 || H <= 0) return fail(c, DRR_E_INVALID, "drr_postprocess: bad arguments");
    if ((post_flags & DRR_POST_NOISE) && !photon_prob) return fail(c, DRR_E_INVALID, "drr_postprocess: noise needs photon_prob");
    if (post_flags & DRR_POST_COLLECTED) return fail(c, DRR_E_INVALID, "drr_postprocess: collected energy is only available in drr_project");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t npix = (size_t)W * H, total = npix * n_views;
    int rc;
    float* d_img = images;
    const float* d_pp = photon_prob;
    if (mem_kind == DRR_MEM_HOST) {
        if ((rc = ensure(c, (void**)&c->d_intensity, &c->int_cap, sizeof(float) * total))) return rc;
        CU(c, cudaMemcpyAsync(c->d_intensity, images, sizeof(float) * total, cudaMemcpyHostToDevice, s));
        d_img = c->d_intensity;
        if (photon_prob) {
            if ((rc = ensure(c, (void**)&c->d_pprob, &c->pp_cap, sizeof(float) * total))) return rc;
            CU(c, cudaMemcpyAsync(c->d_pprob, photon_prob, sizeof(float) * total, cudaMemcpyHostToDevice, s));
            d_pp = c->d_pprob;
        }
    }
    if (post_flags & DRR_POST_NOISE) {
        if ((rc = ensure(c, (void**)&c->d_scratch, &c->scratch_cap, sizeof(float) * total))) return rc;
        CU(c, drr_launch_noise(d_img, d_pp, c->d_scratch, W, H, n_views, photon_count, seed, s));
        c->launches += 2;
    }
    if (post_flags & DRR_POST_CLIP) { CU(c, drr_launch_clip(d_img, total, intensity_upper_bound, s)); c->launches += 1; }
    if (post_flags & DRR_POST_NEGLOG) {
        if (c->minmax_cap < n_views) {
            cudaFree(c->d_minmax); cudaFree(c->d_viewsum);
            CU(c, cudaMalloc(&c->d_minmax, sizeof(unsigned) * 2 * n_views));
            CU(c, cudaMalloc(&c->d_viewsum, sizeof(double) * n_views));
            c->minmax_cap = n_views;
        }
        CU(c, drr_launch_neglog(d_img, npix, n_views, c->d_minmax, 0.01f, s));
        c->launches += 2;
    }
    if (mem_kind == DRR_MEM_HOST) CU(c, cudaMemcpyAsync(images, d_img, sizeof(float) * total, cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    return DRR_OK;
}

int drr_last_timing(const drr_ctx* c, float* ms3) {
    if (!c || !ms3) return DRR_E_INVALID;
    ms3[0] = c->last_ms[0]; ms3[1] = c->last_ms[1]; ms3[2] = c->last_ms[2];
    return DRR_OK;
}

int drr_last_sample_count(const drr_ctx* c, unsigned long long* samples) {
    if (!c || !samples) return DRR_E_INVALID;
    *samples = c->last_samples[0];
    return DRR_OK;
}

int drr_last_window_samples(const drr_ctx* c, unsigned long long* samples) {
    if (!c || !samples) return DRR_E_INVALID;
    *samples = c->last_samples[1];
    return DRR_OK;
}

int drr_host_alloc(size_t bytes, void** out) {
    if (!out || bytes == 0) return fail(nullptr, DRR_E_INVALID, "drr_host_alloc: bad arguments");
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { *out = nullptr; return fail(nullptr, DRR_E_NOMEM, "drr_host_alloc: %s", cudaGetErrorString(e)); }
    return DRR_OK;
}

int drr_host_free(void* p) {
    if (p) cudaFreeHost(p);
    return DRR_OK;
}

int drr_launch_count(const drr_ctx* c, unsigned long long* launches) {
    if (!c || !launches) return DRR_E_INVALID;
    *launches = c->launches;
    return DRR_OK;
}

}  // extern "C"
