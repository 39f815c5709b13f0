/* drr_b200.h -- C ABI of libdrr_b200.so, the B200-native (sm_100a) DRR projection library.
 *
 * This is the drop-in boundary for the projection path of arcadelab/deepdrr.  The reference has no
 * stable FFI for this path: its "ABI" is the positional CuPy RawKernel launch of `projectKernel`
 * (deepdrr/projector/projector.py:718-774, prototype deepdrr/projector/project_kernel.cu:136-181)
 * plus the texture/array set-up in `Projector.initialize` (projector.py:1395-1717).  Each entry point
 * below names the reference code it replaces.  Plain pointers and sizes only; no torch / CuPy types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative DRR_E_* code; drr_last_error() gives text.
 *   - one handle per GPU; a handle is not thread-safe; all work of a handle runs on its own stream
 *     unless a stream is passed to drr_set_stream().
 *   - "mem kind": DRR_MEM_HOST (pageable or pinned host pointer) or DRR_MEM_DEVICE (device pointer
 *     on the handle's GPU, e.g. torch.Tensor.data_ptr()).
 *   - volumes arrive as the reference holds them on the host: float32 density [Ni][Nj][Nk] and
 *     uint8 labels [Ni][Nj][Nk] already remapped to the global material index
 *     (projector.py:1499-1509), C order (k fastest).
 *   - images leave in the orientation Projector.project returns: [view][H][W] float32
 *     (projector.py:786-792 does the (W,H)->(H,W) swap on the host; here the kernel writes it).
 */
#ifndef DRR_B200_H
#define DRR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRR_MAX_VOLUMES 8
#define DRR_MAX_MATERIALS 16

#define DRR_OK 0
#define DRR_E_INVALID (-1)   /* bad argument / unsupported configuration  -> ValueError   */
#define DRR_E_CUDA (-2)      /* CUDA runtime failure                       -> RuntimeError */
#define DRR_E_STATE (-3)     /* call order (e.g. project before volumes)   -> RuntimeError */
#define DRR_E_NOMEM (-4)     /* device allocation failed                   -> MemoryError  */

#define DRR_MEM_HOST 0
#define DRR_MEM_DEVICE 1

/* density sampling back-end of the ray march (all three give the reference's arithmetic) */
#define DRR_SAMPLER_ALU 0     /* texture-unit arithmetic emulated on the SIMT pipes from cell records */
#define DRR_SAMPLER_TEX 1     /* hardware tex3D fetch (the reference's own path, projector.py:116-257) */
#define DRR_SAMPLER_HYBRID 2  /* warps split between the two so TEX units and FMA pipes run together */

/* output post-processing flags for drr_project (projector.py:691-702) */
#define DRR_POST_NEGLOG 1u      /* utils.neglog: -log(I + min(I) + 0.01), min-max to [0,1]          */
#define DRR_POST_NOISE 2u       /* analytic_generators.add_noise: Poisson shot noise + 3x3 blur      */
#define DRR_POST_CLIP 4u        /* np.clip(images, None, intensity_upper_bound)                      */
#define DRR_POST_COLLECTED 8u   /* _calculate_collected_energy_per_pixel (projector.py:833-853)      */

typedef struct drr_ctx drr_ctx;

/* Replaces: cupy.cuda.Device(id).__enter__() + EGL/context set-up (projector.py:1410-1421). */
int drr_create(int device_id, drr_ctx** out);
/* Replaces: Projector.free (projector.py:1719-1764).  Frees every device allocation of the handle. */
int drr_destroy(drr_ctx* ctx);
/* Text of the last error of this handle (or of the failed drr_create when ctx == NULL). */
const char* drr_last_error(const drr_ctx* ctx);
/* Run the handle's work on a caller-owned cudaStream_t (0 = the handle's own stream). */
int drr_set_stream(drr_ctx* ctx, void* cuda_stream);

/* Replaces: energies_gpu / pdf_gpu / absorption_coef_table_gpu uploads (projector.py:1659-1686).
 * energies_keV[n_bins], pdf[n_bins], mu_over_rho[n_bins * n_materials] (index bin * M + m). */
int drr_set_spectrum(drr_ctx* ctx, int n_bins, int n_materials, const float* energies_keV, const float* pdf,
                     const float* mu_over_rho);

/* Replaces: per-volume create_cuda_texture for density (linear) and labels (point)
 * (projector.py:1463-1546, 116-257).  Builds on the device, from one upload of the raw arrays:
 * the [k][j][i] density/label arrays, the 3-D CUDA array + texture object, and the per-cell records
 * the ALU sampler marches over.  Returns the volume index in *vol_id (order = kernel volume order).
 * flags: bit0 = skip cell records (TEX-only handle), bit1 = skip texture (ALU-only handle), bit2 = skip only the
 * 32 B / cell filter-coefficient records (done automatically when they do not fit in device memory: the volume is then
 * sampled by the texture unit alone). */
int drr_add_volume(drr_ctx* ctx, const float* density, const uint8_t* labels, int ni, int nj, int nk, int mem_kind,
                   unsigned flags, int* vol_id);
/* Same, from a Hounsfield-unit volume: HU -> density (vol/volume.py:338-351) and threshold segmentation
 * air <= -800 < soft tissue <= 350 < bone (load_dicom.py:132-143, vol/volume.py:955-992) run on the device, so the
 * host never materialises density / label arrays.  labels_air_soft_bone = global material indices of the three classes. */
int drr_add_volume_hu(drr_ctx* ctx, const float* hu, int ni, int nj, int nk, int mem_kind, const int* labels_air_soft_bone,
                      unsigned flags, int* vol_id);
/* Drop all volumes (keeps spectrum).  Replaces the texture teardown in Projector.free. */
int drr_clear_volumes(drr_ctx* ctx);

/* Replaces: priorities_gpu / volume_enabled_gpu (projector.py:1606-1611, 674-675). */
int drr_set_priorities(drr_ctx* ctx, const int* priority, const int* enabled, int n_volumes);

/* Ray-march options: step (world mm; projector.py:723), attenuate_outside_volume + air_index
 * (projector.py:554-568, -D ATTENUATE_OUTSIDE_VOLUME / AIR_INDEX), sampler = DRR_SAMPLER_*. */
int drr_set_march(drr_ctx* ctx, float step, int attenuate_outside_volume, int air_index, int sampler);

/* Tuning knobs; results stay within the parity tolerance whatever they are set to.
 *   DRR_TUNE_TEX_EIGHTHS    DRR_SAMPLER_HYBRID: how many of every 8 consecutive steps of a uniform segment fetch
 *                           density through the texture unit (the rest emulate it on the FMA pipes); 0, 3, 4 (default),
 *                           5 or 8 (1..2 run as 3, 6..7 as 5).
 *   DRR_TUNE_KERNEL_VARIANT single-volume march: 0 = warp-cooperative shared-memory staging (default),
 *                           1 = per-ray register cell cache.
 *   DRR_TUNE_PIPELINE       k > 0 (default 1): a batch of >= 4 views bound for host memory is projected in two pieces, the
 *                           last k views (at most half) apart, so that the device-to-host copy of the first piece runs
 *                           under the march of the second; 0: one piece.
 *   DRR_TUNE_LANE_QUADS     single-volume lock-step march: 1 = the four lanes the texture unit filters together walk a 2 x 2
 *                           block of pixels, 0 = a 4 x 1 run, 2 (default) = the library's choice (currently 1).
 *   DRR_TUNE_RAYS_PER_LANE  single-volume lock-step march: rays a lane walks through each staged box, 1 or 2; 0 (default) = chosen
 *                           per batch from the ray spacing (two while an 8 x 4 pixel tile spans at most ~2 voxels). */
#define DRR_TUNE_TEX_EIGHTHS 0
#define DRR_TUNE_KERNEL_VARIANT 1
#define DRR_TUNE_PIPELINE 2
#define DRR_TUNE_LANE_QUADS 3
#define DRR_TUNE_RAYS_PER_LANE 4
int drr_set_tuning(drr_ctx* ctx, int key, int value);

/* Mesh inputs of projectKernel (project_kernel.cu:172-177, 363-375, 498-517, 569-579), per view,
 * device or host pointers; NULL disables.  Produced by drr_mesh_* (ray-triangle) or by the caller.
 *   hit_alphas  [n_views][layers][H*W][max_hits] f32, hit_facing same shape i8,
 *   layer_valid [layers] i8, additive [n_views][layers][n_mesh_mats][H*W][2] f32, mesh_mats [n_mesh_mats]. */
int drr_set_mesh_buffers(drr_ctx* ctx, int layers, int max_hits, const float* hit_alphas, const int8_t* hit_facing,
                         const int8_t* layer_valid, const float* additive, const int* mesh_mats, int n_mesh_mats,
                         int mem_kind);

/* Replaces: the pyrender scene + OpenGL renderer set-up (projector.py:1564-1598) and, per projection, the whole
 * _render_mesh path (projector.py:1055-1330: GL additive passes, dual depth peeling, kernelReorder / kernelTide /
 * kernelReorder2, MESH_SUB passes, GL<->CUDA copies) by CUDA ray-triangle intersection.
 *   n_prims primitives; primitive p owns triangles [tri_offsets[p], tri_offsets[p+1]) of `vertices`
 *   ([n_tris][3 vertices][xyz], mesh-local coordinates, outward normals counter-clockwise);
 *   material[p] = global material index, density[p] (g/cm^3), flags[p] bit0 additive / bit1 subtractive
 *   (pyrenderdrr/material.py:41-42), layer[p]; mesh_layers, max_mesh_hits as in Projector (projector.py:419-420).
 * n_prims == 0 removes all meshes. */
int drr_set_meshes(drr_ctx* ctx, int n_prims, const int* tri_offsets, const float* vertices, const int* material,
                   const float* density, const uint8_t* flags, const int* layer, int mesh_layers, int max_mesh_hits);
/* Per batch, before drr_project: world_from_mesh [n_views][n_prims][12] (3x4 row-major), source_world
 * [n_views][3], far_limit = 2 * source_to_detector_distance (projector.py:1227).  Replaces
 * _setup_pyrender_scene (projector.py:855-880). */
int drr_set_mesh_poses(drr_ctx* ctx, int n_views, const float* world_from_mesh, const float* source_world,
                       float far_limit);
/* Replaces: kernelTide (peel_postprocess_kernel.cu:15-177) from its cut-off step on: cleans n_rays hit
 * lists of n (<= 128) slots (distance, facing: +1 entry / -1 exit / 0 empty) in place. */
int drr_mesh_clean_hits(drr_ctx* ctx, float* ts, int8_t* facing, int n_rays, int n, float far_limit, int mem_kind);
/* Replaces: Projector.project_seg / project_hits / project_travel (projector.py:945-1053) for one view of the
 * primitives selected by `select` [n_prims] (u8, non-zero = the primitive carries the requested tag):
 *   DRR_MESH_QUERY_HITS   -> out f32 [H*W][max_mesh_hits]: cleaned (entry, exit, ...) distances, +inf padded; every
 *                            selected primitive counts as subtractive, whatever its layer (renderer.py:312-324,
 *                            force_all_subtract), then kernelReorder/kernelTide (projector.py:1144-1250);
 *   DRR_MESH_QUERY_TRAVEL -> out f32 [H*W]: path length (mm) inside the selected additive layer-0 primitives
 *                            (density pass with density_override=1, projector.py:1027-1053), 0 where the
 *                            entry/exit count does not balance (|G| > 0.01) or the sum is negative;
 *   DRR_MESH_QUERY_SEG    -> out u8 [H*W]: 255 where any triangle of a selected primitive covers the pixel
 *                            (segmentation.frag with GL_MAX blending, renderer.py:326-333, 437-446).
 * Poses come from the last drr_set_mesh_poses (its first view); world_from_index is that view's 3x3.
 * `out` is host or device memory (mem_kind). */
#define DRR_MESH_QUERY_HITS 0
#define DRR_MESH_QUERY_TRAVEL 1
#define DRR_MESH_QUERY_SEG 2
int drr_mesh_query(drr_ctx* ctx, int mode, int width, int height, const float* world_from_index, const uint8_t* select,
                   void* out, int mem_kind);

/* Monte Carlo scatter (north_star kernel 3).  The reference has no scatter kernel any more (projector.py:530-531
 * raises); these entry points take its MC-GPU data tables (mcgpu_mfp_data.py, mcgpu_rita_samplers.py,
 * mcgpu_compton_data.py, mcgpu_density.py) and one volume added with drr_add_volume.
 *   mfp [n_mat][n_e][5] = Rayleigh, Compton, photoelectric, total mean free paths in mm at nominal density and the
 *   Rayleigh max cumulative probability, on the energy grid e0 + i*de (eV); rita [n_mat][128][4] = x^2, P, A, B;
 *   compton [n_mat][30][3] = shell electrons, ionisation energy (eV), J0; rho_nom [n_mat];
 *   mat_of_label [M]: global material index -> table material; rho_max_of_label [M]: largest voxel density per label. */
int drr_set_scatter_tables(drr_ctx* ctx, int n_mat, int n_e, float e0, float de, const float* mfp, const float* rita,
                           const float* compton, const int* nshell, const float* rho_nom, const int* mat_of_label,
                           const float* rho_max_of_label);
/* Simulates photons [photon_offset, photon_offset + n_photons) of the stream `seed` for one view and returns the
 * scatter tally: out_tally [H][W] uint64, energy x solid-angle weight in units of 2^-16 eV (exact integer sums, so
 * any split of the photon range over calls or GPUs adds up bit-identically; reduce across GPUs with ncclAllReduce).
 *   index_from_world: 3x4 K[R|t] divided by the source-to-detector distance (w == 1 on the detector plane);
 *   ijk_from_world [V][12]: one 3x4 per volume added (a point inside several volumes belongs to the one drr_set_priorities
 *   ranks first, as in the ray march; space between the volumes is vacuum).
 *   out_counters (host, may be NULL): energy bookkeeping, eV x weight: emitted, missed the volume, absorbed, left
 *   unscattered, scattered & detected, scattered & missed the detector; then #Rayleigh and #Compton events. */
int drr_scatter(drr_ctx* ctx, unsigned long long n_photons, unsigned long long photon_offset, uint64_t seed, int W, int H,
                const float* world_from_index, const float* index_from_world, const float* source_world,
                const float* ijk_from_world, unsigned long long* out_tally, double* out_counters, int out_mem_kind);

/* Replaces: the per-view loop of Projector.project -> _render_single (projector.py:679-685, 709-800):
 * _update_object_locations uploads (802-831), the projectKernel launch (770-774), the two D2H copies
 * and swapaxes (786-792) and the host post-processing (691-702), for a whole batch of views.
 *   world_from_index [n_views][9], source_ijk [n_views][V][3], ijk_from_world [n_views][V][12]: host f32.
 *   out_intensity / out_photon_prob: [n_views][H][W] f32 (out_photon_prob may be NULL);
 *   out_area: [n_views][M][H][W] f32 per-material area densities in g/cm^2 (may be NULL).
 *   post_flags: DRR_POST_*; photon_count / intensity_upper_bound / seed used by the flags that need them. */
int drr_project(drr_ctx* ctx, int n_views, int W, int H, const float* world_from_index, const float* source_ijk,
                const float* ijk_from_world, float max_ray_length, unsigned post_flags, float photon_count,
                float intensity_upper_bound, float pixel_area_mm2, uint64_t seed, float* out_intensity,
                float* out_photon_prob, float* out_area, int out_mem_kind);

/* Replaces: the host post-processing of Projector.project (projector.py:691-702) for images that were
 * modified after drr_project (e.g. primary + scatter): noise, clip, neglog on [n_views][H][W] in place. */
int drr_postprocess(drr_ctx* ctx, float* images, const float* photon_prob, int n_views, int W, int H, unsigned post_flags,
                    float photon_count, float intensity_upper_bound, uint64_t seed, int mem_kind);

/* Kernel-only timing of the last drr_project (CUDA events on the handle's stream), ms:
 * [0] ray march, [1] spectral/post kernels, [2] whole call incl. copies. */
int drr_last_timing(const drr_ctx* ctx, float* ms3);
/* Sum over the last batch of max(num_steps, 0) (x volumes traced) -- SURVEY.md 8(d) "S_view". */
int drr_last_sample_count(const drr_ctx* ctx, unsigned long long* samples);
/* Of those, the steps that fell inside a volume's [lo, hi] window (the samples that fetch density); filled by the
 * single-volume lock-step kernel, 0 otherwise.  bench.py's gather rate is quoted on this figure. */
int drr_last_window_samples(const drr_ctx* ctx, unsigned long long* samples);
/* Page-locked host memory for image outputs, so that the device-to-host copy at the end of drr_project runs at full PCIe
 * rate.  Replaces: the pageable ndarray cupy's `.get()` returns at projector.py:786-787. */
int drr_host_alloc(size_t bytes, void** out);
int drr_host_free(void* p);
/* Number of kernels this library launched since the handle was created. */
int drr_launch_count(const drr_ctx* ctx, unsigned long long* launches);
/* Block until all work of the handle has finished. */
int drr_synchronize(drr_ctx* ctx);

/* Library / build info: "drr_b200 <version> sm_100a". */
const char* drr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DRR_B200_H */
