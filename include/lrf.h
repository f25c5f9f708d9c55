/* lrf.h — C ABI of the B200-native LiteRF ray-march library (liblrf_b200.so).
 *
 * One entry point per member of the reference's tracer class on the hot path.  A reference
 * maintainer binds these from a replacement example_tracer.{h,cpp} (shipped under
 * literf-radiancefields_b200/host/example_tracer/) so that main.cpp compiles
 * unmodified.  "ref:" cites the reference interface each function replaces, relative to the
 * reference repository root.
 *
 * Conventions
 *   - Every function returns 0 on success, non-zero on failure; lrf_last_error() gives the text.
 *     (The reference has no error convention: all methods are void and failures are crashes.)
 *   - There is NO CPU fallback: lrf_create fails if no CUDA device / sm_100 kernel image is usable.
 *   - Grid host layout is the reference's: gs^3 cells x 13 floats, AoS, x fastest
 *     (ref: example_tracer/example_tracer.h:25-32, example_tracer.cpp:62-65).
 *   - Boxes are unit voxels of the (gs-1)^3 lattice, named by integer coordinates; the position
 *     in the list is the box id (the octree EntityID used as tie-break, ref:
 *     external/Octree/octree.h:856-860).  gs must be a power of two (every size the reference
 *     uses: 16..128, main.cpp:49,125-133) so that box planes are exact in fp32.
 *   - Matrices are 16 floats, column-major, ALREADY INVERTED on the host (the reference inverts in
 *     its setters, example_tracer.h:214-215, with LiteMath::inverse4x4).
 *   - Calls are synchronous unless the name ends in _async; a context is not re-entrant.
 */
#ifndef LRF_H
#define LRF_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lrf_ctx lrf_ctx;

/* ---- lifetime -------------------------------------------------------------------------------- */
/* ref: RayMarcherExample::RayMarcherExample(), example_tracer.h:43-49; CommitDeviceData() :223 */
int lrf_create(int device, lrf_ctx **out);
/* ref: ~RayMarcherExample (main.cpp:143) */
int lrf_destroy(lrf_ctx *h);
const char *lrf_last_error(void);
/* stream-sync; ref: UpdateMembersPlainData(), example_tracer.h:224 (the natural sync point) */
int lrf_sync(lrf_ctx *h);

/* ---- grid / boxes ---------------------------------------------------------------------------- */
/* ref: InitGrid(float), example_tracer.h:51-91 — density 0.01, other 12 floats 0.1, dense (gs-1)^3 boxes */
int lrf_init_grid(lrf_ctx *h, int gs);
/* no reference counterpart: the seeded synthetic grids of the benchmark/test workloads, generated in place on
 * the device.  mode 0 = LCG grid (density_scale D: 20 "opaque", 2 "thin"), 1 = LCG x Gaussian blob, 2 = sphere-signed
 * density (positive inside r = 0.288, negative outside: the ~10 % occupancy scene of BASELINE config 4). */
int lrf_fill_synthetic(lrf_ctx *h, float density_scale, int mode);
/* ref: InitGrad(), example_tracer.h:93-96 (allocates; contents zero) */
int lrf_init_grad(lrf_ctx *h);
/* ref: zeroGrad(), example_tracer.h:104-109 */
int lrf_zero_grad(lrf_ctx *h);
/* ref: SetBoundingBox, example_tracer.h:98-102 (only bb.max.x-bb.min.x is used: example_tracer.cpp:649,692) */
int lrf_set_bbox(lrf_ctx *h, const float mn[3], const float mx[3]);
int lrf_grid_size(lrf_ctx *h);
/* ref: public member `grid` (raw model.dat dump/load, main.cpp:54-56,139-141) */
int lrf_grid_upload(lrf_ctx *h, const float *cells13);
int lrf_grid_download(lrf_ctx *h, float *cells13);
/* ref: public member `grid_d` */
int lrf_grad_upload(lrf_ctx *h, const float *cells13);
int lrf_grad_download(lrf_ctx *h, float *cells13);
/* ref: public members V, S (example_tracer.h:248-249); which: 0 = V, 1 = S */
int lrf_adam_state_upload(lrf_ctx *h, int which, const float *cells13);
int lrf_adam_state_download(lrf_ctx *h, int which, float *cells13);
/* ref: public member `boxes` (+ octree rebuild, example_tracer.h:87).  xyz = n x 3 lattice coords in id order.
 * lrf_set_boxes_dense restores the InitGrid list. */
int lrf_set_boxes(lrf_ctx *h, const int32_t *xyz, size_t n);
int lrf_set_boxes_dense(lrf_ctx *h);
/* ref: boxes.size() (main.cpp:136) */
size_t lrf_box_count(lrf_ctx *h);
int lrf_get_boxes(lrf_ctx *h, int32_t *xyz);

/* ---- camera ---------------------------------------------------------------------------------- */
/* ref: SetWorldViewMatrix / SetWorldViewMProjatrix, example_tracer.h:214-215 (members m_worldViewInv, m_worldViewProjInv) */
int lrf_set_view_inv(lrf_ctx *h, const float m16[16]);
int lrf_set_proj_inv(lrf_ctx *h, const float m16[16]);

/* ---- forward --------------------------------------------------------------------------------- */
/* ref: kernel2D_RayMarch(uint32_t*,uint32_t,uint32_t,bool), example_tracer.cpp:627-663.
 * out = width*height packed pixels (R | G<<8 | B<<16 | A<<24, row 0 = bottom), HOST memory. */
int lrf_render(lrf_ctx *h, uint32_t *out_host, uint32_t width, uint32_t height, int greyscale);
/* same, output left in DEVICE memory (out_dev), asynchronous on the context stream */
int lrf_render_dev_async(lrf_ctx *h, uint32_t *out_dev, uint32_t width, uint32_t height, int greyscale);
/* n_views frames in one call (ref: the caller's loop over views, main.cpp:87-119 — SetWorldViewMatrix :104, then L1Loss :114, which
 * calls kernel2D_RayMarch, example_tracer.cpp:603-625 — once per view): frame v of out_dev (width*height pixels each, DEVICE memory) is rendered with the inverse view matrix
 * view_inv_host[16 v ..] (HOST memory, copied before the call returns); projection, grid and options are the context's and its own
 * view matrix is left unchanged.  One kernel launch for all views (dense, sparse and greyscale alike; the opt-in staged kernel: one per view). */
int lrf_render_views_dev_async(lrf_ctx *h, uint32_t n_views, const float *view_inv_host, uint32_t *out_dev, uint32_t width, uint32_t height,
                               int greyscale);
/* same with the frames delivered to HOST memory (frame v at out_host + v*width*height), synchronous like kernel2D_RayMarch: complete on
 * return.  Up to four batches; a finished batch is copied out while the next one renders (pin out_host for the overlap). */
int lrf_render_views(lrf_ctx *h, uint32_t n_views, const float *view_inv_host, uint32_t *out_host, uint32_t width, uint32_t height, int greyscale);
/* rows [y0, y1) of the frame only, written into the full-frame buffer out_dev: one frame split over several GPUs in row bands
 * (ref: kernel2D_RayMarch parallelises the rows of a frame, example_tracer.cpp:630) */
int lrf_render_rows_dev_async(lrf_ctx *h, uint32_t *out_dev, uint32_t width, uint32_t height, uint32_t y0, uint32_t y1, int greyscale);
/* lrf_render / lrf_render_dev_async use the thread-per-ray gather kernel; with option "staged_render" = 1 dense colour grids of
 * 8..512 nodes per axis are rendered by the TMA-staged kernel (lrf_staged.cuh) instead (fewer L1 wavefronts, more instructions: slower
 * on the bench workload).  Parity tap of the staged kernel: the frame and, per pixel, the number of samples it shaded before the
 * reference's `throughput < 0.1` break (ref example_tracer.cpp:131) — 0xffffffff where the pixel was handed to the exact
 * kernel (break too close to call on the approximate throughput, tie events, axis-aligned rays).  HOST arrays, may be NULL. */
int lrf_render_staged_debug(lrf_ctx *h, uint32_t *out_host, uint32_t *n_shaded_host, uint32_t width, uint32_t height);
/* Dense colour frames are rendered by the register-window march (lrf_lean.cuh; option "lean_render" = 0 selects the older gather
 * kernel).  Its parity tap: the frame, per pixel the number of samples shaded before the reference's `throughput < 0.1` break
 * (ref example_tracer.cpp:131; 0xffffffff where the pixel took the generic fallback) and the final transmittance (bits comparable
 * with lrf_render_debug's).  HOST arrays, may be NULL.  Dense box sets only. */
int lrf_render_lean_debug(lrf_ctx *h, uint32_t *out_host, uint32_t *n_shaded_host, float *thr_host, uint32_t width, uint32_t height);
/* counters since the last reset: [0] rays re-rendered by the exact kernel because the break was too close to call,
 * [1] other rays sent there, [2] pixel tiles whose footprint did not fit the staged box, [3] 1 if the last render was staged */
int lrf_render_stats(lrf_ctx *h, uint32_t out4[4], int reset);
/* parity taps (no reference counterpart; what RayIntersectedAll + RayGridIntersection compute per
 * ray, example_tracer.cpp:641-651).  Host arrays, any may be NULL: n_hit[w*h], n_shaded[w*h],
 * rgbtc[w*h*5] = {r,g,b,T,cauchy}, vox_xyz[w*h*keep*3] first `keep` hit voxels (-1 padded). */
int lrf_render_debug(lrf_ctx *h, uint32_t width, uint32_t height, int greyscale, uint32_t *n_hit, uint32_t *n_shaded,
                     float *rgbtc, int32_t *vox_xyz, uint32_t keep);
/* packed rays (n x 3 origins, n x 3 directions, HOST memory): same taps */
int lrf_trace_rays(lrf_ctx *h, size_t n, const float *o, const float *d, int greyscale, uint32_t *n_hit, uint32_t *n_shaded,
                   float *rgbtc, int32_t *vox_xyz, uint32_t keep);

/* ---- backward -------------------------------------------------------------------------------- */
/* ref: kernel2D_RayMarchGrad(uint32_t,uint32_t,float4*,bool), example_tracer.cpp:665-718 (+ RayGridLossGrad
 * :270-279, the Enzyme call).  ref_rgba = width*height float4 (stbi layout, alpha ignored), HOST memory;
 * it is fully consumed before the call returns.  Accumulates (+=) into the gradient grid. */
int lrf_render_grad(lrf_ctx *h, uint32_t width, uint32_t height, const float *ref_rgba_host, int greyscale);
/* device-resident variant: ref image already in DEVICE memory, asynchronous */
int lrf_render_grad_dev_async(lrf_ctx *h, uint32_t width, uint32_t height, const float *ref_rgba_dev, int greyscale);
/* packed-ray batch (BASELINE config 3): n rays, o/d n x 3, ref n x 4, all DEVICE memory, asynchronous */
int lrf_rays_grad_dev_async(lrf_ctx *h, size_t n, const float *o_dev, const float *d_dev, const float *ref_rgba_dev,
                            int greyscale);
/* host-memory convenience wrapper of the above (synchronous) */
int lrf_rays_grad(lrf_ctx *h, size_t n, const float *o, const float *d, const float *ref_rgba, int greyscale);
/* device-side ray generation for a batch of (view, pixel) picks: view_inv = n_views x 16 floats (DEVICE),
 * picks = n x 2 uint32 {view, pixel index y*w+x} (DEVICE); writes o_dev/d_dev.  ref: example_tracer.cpp:635-639 */
int lrf_gen_rays_dev_async(lrf_ctx *h, size_t n, const float *view_inv_dev, const uint32_t *picks_dev, uint32_t width,
                           uint32_t height, float *o_dev, float *d_dev);

/* ---- TV / optimiser -------------------------------------------------------------------------- */
/* ref: kernel2D_TVGrad(), example_tracer.cpp:737-749 (+ TVRegularisationGrad :250-256) */
int lrf_tv_grad(lrf_ctx *h);
/* ref: optimizerInit(), example_tracer.h:111-126 */
int lrf_adam_init(lrf_ctx *h);
/* ref: optimizerStep(int) :128-163 (which = 0), optimizerStepDensity :165-183 (1), optimizerStepSH :185-206 (2) */
int lrf_adam_step(lrf_ctx *h, int iter, int which);
/* fused per-cell pass for training loops: TV gradient + Adam (all channels) + gradient zeroing for the
 * next iteration in one call (same results as lrf_tv_grad; lrf_adam_step(iter,0); lrf_zero_grad).  Asynchronous. */
int lrf_train_percell_async(lrf_ctx *h, int iter);
int lrf_set_adam_hparams(lrf_ctx *h, float lr, float beta1, float beta2, float eps);
/* the same iteration with the TV gradient running BESIDE the ray backward (second stream; TV reads parameters only and both
 * scatter with red.add): lrf_rays_grad_tv_dev_async = lrf_rays_grad_dev_async || kernel2D_TVGrad (own z-slab only when a peer
 * group is attached), joined on the context stream; lrf_adam_zero_async = optimizerStep(iter) + zeroGrad.  Asynchronous. */
int lrf_rays_grad_tv_dev_async(lrf_ctx *h, size_t n, const float *o_dev, const float *d_dev, const float *ref_rgba_dev, int greyscale);
int lrf_adam_zero_async(lrf_ctx *h, int iter);

/* ---- multi-GPU exchange over NVLink peer memory (no reference counterpart: the reference is one process) ----
 * One process per GPU, a grid replica each.  After the backward, rank r owns z-slab r of the planes: it adds the TV
 * gradient of its slab, reads the gradient of EVERY rank for that slab (peer loads), applies Adam with its local moments
 * and stores the new parameters into EVERY replica (peer stores) — reduce-scatter + TV + Adam + all-gather in one pass.
 * The caller orders the GPUs: a barrier between the backward and lrf_train_percell_peer_async, another one after it, then
 * lrf_zero_grad_async.  In this mode the Adam moments of rank r are current for slab r only.
 * gridSize must be a multiple of 8 * world; world <= 8. */
/* 64-byte CUDA IPC handle of the parameter (which = 0) or gradient (1) planes or the barrier flags (2), to be sent to the other ranks */
int lrf_ipc_export(lrf_ctx *h, int which, void *handle64);
/* param_handles / grad_handles: world x 64 bytes, rank-ordered (the own rank's entry is ignored) */
int lrf_peer_attach_ipc(lrf_ctx *h, int world, int rank, const void *param_handles, const void *grad_handles);
/* the same for contexts living in THIS process (one per device, or several on one device in tests) */
int lrf_peer_attach_local(lrf_ctx *h, int world, int rank, lrf_ctx *const *ctxs);
int lrf_peer_detach(lrf_ctx *h);
/* the barrier the exchange needs, over peer memory instead of a collective: every rank stores an epoch into every rank's flag
 * array and waits for all slots of its own (release / acquire at system scope; ~10 us).  lrf_ipc_export(h, 2, ..) exports the
 * flag array; lrf_peer_attach_flags_ipc maps the group's (world x 64 bytes, after lrf_peer_attach_ipc; a group attached with
 * lrf_peer_attach_local needs no call).  A peer that never arrives trips a 2 s time-out instead of hanging the GPU. */
int lrf_peer_attach_flags_ipc(lrf_ctx *h, const void *flag_handles);
int lrf_peer_barrier_async(lrf_ctx *h);
int lrf_peer_barrier_timed_out(lrf_ctx *h);
int lrf_tv_grad_slab_async(lrf_ctx *h);              /* kernel2D_TVGrad restricted to the own slab */
/* The same exchange through the NVSwitch — NVLink multicast objects: multimem.ld_reduce returns the sum of every replica's gradient
 * in one reduced response, multimem.st stores the new parameters into every replica with one store (half the NVLink volume of the
 * point-to-point pass in each direction).  After the four steps below lrf_adam_peer_async / lrf_train_percell_peer_async take this
 * path; lrf_peer_detach (or destroying the grid) leaves the group.  While attached the parameter and gradient planes live in
 * cuMemCreate allocations (lrf_ipc_export of them is not available; the flag barrier is unchanged: lrf_ipc_export(h, 2, ..) +
 * lrf_peer_attach_flags_ipc).  The switch sums in its own order: results can differ from the rank-order sum in the last bit, the
 * replicas stay identical to each other.
 *   1. every rank:  lrf_mc_prepare(h, world, rank)
 *   2. one rank:    lrf_mc_create(h, &fd)  -> pass fd to the other processes (unix socket, SCM_RIGHTS);  others: lrf_mc_import(h, fd)
 *   3. every rank:  lrf_mc_add_device(h);  group-wide barrier
 *   4. every rank:  lrf_mc_bind(h);        group-wide barrier
 * lrf_mc_attach_local does all of it for contexts of one process (and wires the barrier flags like lrf_peer_attach_local). */
int lrf_mc_supported(int device);
int lrf_mc_prepare(lrf_ctx *h, int world, int rank);
int lrf_mc_create(lrf_ctx *h, int *fd_out);
int lrf_mc_import(lrf_ctx *h, int fd);
int lrf_mc_add_device(lrf_ctx *h);
int lrf_mc_bind(lrf_ctx *h);
int lrf_mc_attach_local(int world, lrf_ctx *const *ctxs);
int lrf_mc_attached(lrf_ctx *h);
int lrf_adam_peer_async(lrf_ctx *h, int iter);       /* optimizerStep(iter) on the own slab, gradient summed over the ranks */
int lrf_train_percell_peer_async(lrf_ctx *h, int iter);   /* the two above */
int lrf_zero_grad_async(lrf_ctx *h);                 /* zeroGrad without the stream sync */

/* ---- level transition ------------------------------------------------------------------------ */
/* ref: UpsampleGrid(), example_tracer.cpp:281-323 */
int lrf_upsample_grid(lrf_ctx *h);
/* ref: UpsampleOctree(), example_tracer.cpp:325-447 */
int lrf_upsample_boxes(lrf_ctx *h);
/* ref: RebuildOctree(), example_tracer.cpp:449-489 */
int lrf_rebuild_boxes(lrf_ctx *h);

/* ---- callers either side of the path ------------------------------------- */
/* ref: the raw model.dat dump / load kept (commented) in main.cpp:54-56,139-141: gs^3 cells x 13 floats, no header.
 * lrf_model_load fails unless the file size equals the current grid's gs^3 * 52 bytes. */
int lrf_model_save(lrf_ctx *h, const char *path);
int lrf_model_load(lrf_ctx *h, const char *path);
/* Resident target images.  ref: main.cpp:72,88-97 — stbi_set_flip_vertically_on_load(true) + stbi_loadf(.., 4): every
 * view's PNG is decoded, flipped, linearised (external/stb_image.h:1856-1876: colour = pow(u8/255.0f, 2.2f), alpha =
 * u8/255.0f) and passed to L1Loss as float4[w*h], then freed.  Here the caller hands over the DECODED 8-bit RGBA rows
 * (file order, top row first; PNG decoding stays on the host) and the flip + linearisation run on the device into
 * slot `view` of a resident array, bit-identical to stbi_loadf's output. */
int lrf_targets_resize(lrf_ctx *h, uint32_t n_views, uint32_t width, uint32_t height);
int lrf_target_ingest_rgba8(lrf_ctx *h, uint32_t view, const uint8_t *rgba8_host, int flip_vertical);
/* already-linear float4 pixels (what stbi_loadf returned), HOST memory */
int lrf_target_upload_rgbaf(lrf_ctx *h, uint32_t view, const float *rgba_host);
int lrf_target_download(lrf_ctx *h, uint32_t view, float *rgba_host);
float *lrf_target_device_ptr(lrf_ctx *h, uint32_t view);
/* kernel2D_RayMarchGrad (example_tracer.cpp:665-718) against resident target `view`, current camera; asynchronous */
int lrf_render_grad_target_async(lrf_ctx *h, uint32_t view, int greyscale);
/* reference pixels for a batch of picks {view, pixel index} (DEVICE, as lrf_gen_rays_dev_async takes them) out of the
 * resident targets -> ref_rgba_dev[n x 4]; asynchronous */
int lrf_gather_targets_dev_async(lrf_ctx *h, size_t n, const uint32_t *picks_dev, float *ref_rgba_dev);
/* ref: LiteImage::SaveBMP(fname, pixels, w, h) as L1Loss calls it (example_tracer.cpp:614; external/LiteMath/
 * Image2d.cpp:488-533): 24-bit BMP, records {(p>>16)&255, (p>>8)&255, p&255}, no row padding, byte-identical file. */
int lrf_save_bmp(const char *path, const uint32_t *pixels_host, int width, int height);
/* kernel2D_RayMarch + the BMP record packing as a device epilogue + file write */
int lrf_render_bmp(lrf_ctx *h, const char *path, uint32_t width, uint32_t height, int greyscale);

/* ---- plumbing -------------------------------------------------------------------------------- */
/* ref: GetExecutionTime(const char*, float[4]), example_tracer.cpp:759-763.  name in {"RayMarch",
 * "RayMarchGrad","TVGrad","Adam"}: out[0] = device ms of the last such call (CUDA events). */
int lrf_last_time_ms(lrf_ctx *h, const char *name, float out[4]);
/* CUDA-event stopwatch on the context stream (the stream every kernel of this context is launched on):
 * lrf_timer_start records an event; lrf_timer_stop_ms records a second one, waits for it and returns the
 * elapsed device time in milliseconds. */
int lrf_timer_start(lrf_ctx *h);
int lrf_timer_stop_ms(lrf_ctx *h, float *ms);
/* device pointers of the parameter / gradient storage in DEVICE layout (density plane [gs^3] followed by
 * colour plane [gs^3 x 12]) for NCCL all-reduce by the multi-GPU harness; n_floats = 13*gs^3 */
float *lrf_param_device_ptr(lrf_ctx *h);
/* a caller that WRITES the parameters through that pointer (or through peer memory) says so here: derived copies are stale */
int lrf_params_touched(lrf_ctx *h);
float *lrf_grad_device_ptr(lrf_ctx *h);
size_t lrf_param_count(lrf_ctx *h);
void *lrf_stream(lrf_ctx *h);              /* cudaStream_t the context launches on */
int lrf_use_stream(lrf_ctx *h, void *cuda_stream);   /* launch on the caller's stream (e.g. torch's current stream) */
int lrf_device_alloc(lrf_ctx *h, size_t bytes, void **out_dev);
int lrf_device_free(lrf_ctx *h, void *dev);
int lrf_host_alloc_pinned(size_t bytes, void **out_host);
int lrf_host_free_pinned(void *host);
int lrf_memcpy_h2d_async(lrf_ctx *h, void *dst_dev, const void *src_host, size_t bytes);
int lrf_memcpy_d2h_async(lrf_ctx *h, void *dst_host, const void *src_dev, size_t bytes);
/* counters the benchmark reports: kernels launched by this context since creation; distinct nodes
 * read by the last lrf_count_touched call */
uint64_t lrf_launch_count(lrf_ctx *h);
/* test / diagnostic switches.  "force_generic_march" = 1: the parity taps (lrf_render_debug, lrf_trace_rays) send every
 * ray down the generic exact march instead of the streamlined one, so that tests can compare the two. */
int lrf_set_option(lrf_ctx *h, const char *name, int value);
int lrf_count_touched_nodes(lrf_ctx *h, uint32_t width, uint32_t height, int greyscale, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
