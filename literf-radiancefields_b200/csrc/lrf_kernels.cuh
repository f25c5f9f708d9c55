// lrf_kernels.cuh — launch interface between the C-ABI layer (lrf_capi.cu) and the sm_100a
// kernels (lrf_kernels.cu).  Plain pointers and sizes only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "lrf_device.cuh"

namespace lrf {

struct DebugTaps  // device pointers, any may be null
{
  uint32_t *n_hit;
  uint32_t *n_shaded;
  float *rgbtc;      // 5 per ray
  int32_t *vox_xyz;  // keep*3 per ray
  uint32_t keep;
  uint8_t *touched;  // [gs^3] node-read marks (lrf_count_touched_nodes)
};


// staged forward render (lrf_staged.cuh): launch description built by the C-ABI layer
struct alignas(64) StagedArgs
{
  CUtensorMap tm[3];        // by march axis A: box {16, 1, W, H} / {16, W, 1, H} / {16, W, H, 1} over {16 floats, x, y, z}
  int W[3], H[3];           // footprint box in nodes by march axis: W along the first lateral axis in memory order, H along the second
  uint32_t plane_bytes[3];  // W * H * 64 (what one TMA delivers)
  uint32_t slot_bytes;      // ring slot stride: the largest plane rounded up to kStSlotAlign
  float err_sigma;          // bound on |weight-form density - reference nested-lerp density|: 16 ulp of the largest |density| in the grid
  uint32_t *diag;           // host-mapped: watchdog record {code, block, thread, a, b, c} written before a stuck wait traps the kernel
  uint32_t *stats;          // [0] rays sent to the exact kernel because the break was ambiguous, [1] other rays sent there (ties, entry on a
                            // plane, axis-aligned / unfit lanes), [2] CTAs whose footprint did not fit the box
};

// forward, staged: bake the 64-byte node records from the planes; render through the TMA-staged kernel (nw = 4 or 8 consumer warps per
// CTA) followed by the exact kernel over the fallback list; tap_shaded optional (w*h uint32)
void launch_bake_records(const SceneDev &S, float4 *rec, float *max_abs_density, cudaStream_t st);
size_t staged_smem_bytes(const StagedArgs &T);
int launch_render_staged(const SceneDev &S, const StagedArgs &T, int nw, uint32_t *out_dev, uint32_t w, uint32_t h, uint32_t *fb_count,
                         uint32_t *fb_list, uint32_t *tap_shaded, cudaStream_t st);   // returns 0, or a cudaError_t from the attribute call

// forward
// fb_count (1 uint32) / fb_list (w*h uint32): fallback queue for rays that need the generic march; 2 kernel launches
// rows [y0, y1) of the frame only (default: all of it)
void launch_render_views(const SceneDev &S, const float *views_dev, uint32_t n_views, uint32_t *out_dev, uint32_t w, uint32_t h, int grey, uint32_t *fb_count,
                         uint32_t *fb_list, cudaStream_t st, int lean);
void launch_render_image(const SceneDev &S, uint32_t *out_dev, uint32_t w, uint32_t h, int grey, uint32_t *fb_count, uint32_t *fb_list,
                         cudaStream_t st, uint32_t y0 = 0, uint32_t y1 = 0xffffffffu, int lean = 0, uint32_t *tap_n = nullptr, float *tap_thr = nullptr);
// lean = 1: dense colour frames by the register-window march (lrf_lean.cuh); tap_n / tap_thr: its per-pixel shaded-sample counts and final
// transmittances (parity tests)
void launch_debug_image(const SceneDev &S, uint32_t w, uint32_t h, int grey, const DebugTaps &taps, int force_generic, cudaStream_t st);
void launch_debug_rays(const SceneDev &S, size_t n, const float *o, const float *d, int grey, const DebugTaps &taps, int force_generic,
                       cudaStream_t st);
// backward
// fb_count / fb_list as for launch_render_image (w*h entries); fb_count == nullptr: the generic kernel for every pixel
void launch_grad_image(const SceneDev &S, uint32_t w, uint32_t h, const float *ref_rgba, int grey, uint32_t *fb_count, uint32_t *fb_list,
                       cudaStream_t st);
// packed rays: warp-per-ray kernel + generic fallback over the queue (fb_count: 1 uint32, fb_list: n uint32); 2 launches
// hit_scratch: grad_rays_hit_bytes(n, gs) bytes where pass A leaves each ray's hit list for pass B, or nullptr (pass B walks again)
size_t grad_rays_hit_bytes(size_t n, int gs, uint32_t *cap_out);
void launch_grad_rays(const SceneDev &S, size_t n, const float *o, const float *d, const float *ref_rgba, int grey, uint32_t *fb_count,
                      uint32_t *fb_list, void *hit_scratch, cudaStream_t st, int parallel_gather = 1);
void launch_gen_rays(const float *proj_inv_host16, size_t n, const float *view_inv_dev, const uint32_t *picks, uint32_t w, uint32_t h,
                     float *o, float *d, cudaStream_t st);
// per-cell
void launch_tv_grad(int gs, const float *dens, float *gdens, float *tmp_v, float *tmp_g, cudaStream_t st);
// z-sweep TV gradient over TMA-fed density planes (lrf_tv.cuh): tm = 3-D tensor map over the density plane array with box {40, 12, 1};
// gs a multiple of 32; nodes with z in [z_begin, z_end); diag = host-mapped watchdog record (may be null)
void launch_tv_sweep(const CUtensorMap &tm, int gs, float *gdens, int z_begin, int z_end, int atomic, uint32_t *diag, cudaStream_t st,
                     int blocks_per_sm = 0);   // > 0: cap the resident blocks per SM (when it shares the GPU with an HBM-bound kernel)
void launch_adam(float *p, float *g, float *V, float *S, size_t off, size_t count, float lr, float beta1, float beta2, float eps,
                 double om_b1, double om_b2, int zero_grad, cudaStream_t st);
void launch_fill(float *p, size_t count, float v, cudaStream_t st);
// inter-GPU barrier over peer memory: flags[r] = rank r's flag array (>= world + 1 uint32; slot [world] unused), epoch strictly increasing
// the exchange through NVLink multicast objects: mc_grad / mc_param = multicast addresses of the gradient / parameter planes of the
// group (multimem.ld_reduce / multimem.st), param = this rank's own (unicast) parameter planes; count a multiple of 4
void launch_adam_mc(const float *mc_grad, float *mc_param, const float *param, float *V, float *S, size_t off, size_t count, float lr, float beta1,
                    float beta2, float eps, double om_b1, double om_b2, cudaStream_t st, int unroll = 4, int blocks_per_sm = 0);
void launch_peer_barrier(int world, int rank, uint32_t *const *flags, uint32_t epoch, uint32_t *timed_out, cudaStream_t st);
// multi-GPU: TV gradient of the nodes in z-planes [z_begin, z_end) (multiples of 8), and the Adam pass over the element range
// [off, off+count) that sums the gradient of all ranks through peer memory and writes the parameters to every replica
void launch_tv_grad_slab(int gs, const float *dens, float *gdens, int z_begin, int z_end, int atomic, int max_blocks, cudaStream_t st);   // gs >= 8 only; max_blocks > 0: persistent grid of that many blocks
void launch_adam_peer(int world, int rank, float *const *grads, float *const *params, float *V, float *S, size_t off, size_t count, float lr,
                      float beta1, float beta2, float eps, double om_b1, double om_b2, cudaStream_t st);
// layout conversion: host AoS [N][13]  <->  device planes dens[N] + col[N][12]
void launch_aos_to_planes(const float *aos, float *planes, size_t N, cudaStream_t st);
void launch_planes_to_aos(const float *planes, float *aos, size_t N, cudaStream_t st);
// callers either side of the path: target ingest (stbi_loadf semantics), batch gather of target pixels, BMP record packing
void launch_ingest_rgba8(const uint8_t *src_rgba8, float *dst_rgba, uint32_t w, uint32_t h, int flip, const float *lut_dev, cudaStream_t st);
void launch_gather_targets(size_t n, const uint32_t *picks, const float *targets, size_t view_stride_px, float *out_rgba, cudaStream_t st);
void launch_pack_bmp(const uint32_t *pix, uint8_t *out, size_t n, cudaStream_t st);
// SURVEY 8(d) synthetic grids generated in place; dense (gs-1)^3 box list in id order
void launch_fill_synthetic(int gs, float D, int mode, float *planes, cudaStream_t st);
void launch_dense_boxes(int n_vox_axis, int32_t *xyz, cudaStream_t st);
// level transition
void launch_upsample_grid(int old_gs, const float *old_planes, float *new_planes, cudaStream_t st);
void launch_box_child_flags(const SceneDev &S_new, const int32_t *old_xyz, size_t n_old, uint32_t *flags, cudaStream_t st);   // 8 per box
void launch_box_keep_flags(const SceneDev &S, const int32_t *xyz, size_t n, uint32_t *flags, cudaStream_t st);                // 1 per box
void launch_box_compact_children(const int32_t *old_xyz, size_t n_old, const uint32_t *flags, const uint32_t *offsets, int32_t *new_xyz,
                                 cudaStream_t st);
void launch_box_compact(const int32_t *xyz, size_t n, const uint32_t *flags, const uint32_t *offsets, int32_t *new_xyz, cudaStream_t st);
void launch_build_box_maps(int n_vox_axis, const int32_t *xyz, size_t n, uint32_t *occ, uint32_t *vox2id, uint32_t *occ8, int nb8, cudaStream_t st);
// dist / tmp: nb8^3 bytes each; dist[b] = Chebyshev distance (bricks, capped) from brick b to the nearest occupied brick
void launch_brick_dist(int nb8, const uint32_t *occ8, uint8_t *dist, uint8_t *tmp, cudaStream_t st);
void launch_count_nonzero(const uint8_t *marks, size_t n, unsigned long long *out, cudaStream_t st);
// exclusive prefix sum of uint32 flags (single-pass decoupled look-back is overkill here; two-level scan)
void launch_exclusive_scan(const uint32_t *in, uint32_t *out, size_t n, uint32_t *block_tmp, cudaStream_t st);
size_t scan_tmp_elems(size_t n);

}  // namespace lrf
