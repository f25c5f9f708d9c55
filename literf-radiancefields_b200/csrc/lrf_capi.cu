// lrf_capi.cu — the C-ABI of include/lrf.h over the sm_100a kernels.  State that the reference keeps in
// std::vector members of RayMarcherExample (example_tracer/example_tracer.h:232-251) lives in HBM here.
//
// HBM layout per context (N = gs^3):
//   params / grad / adam V / adam S : 13*N floats each = [density plane N][colour plane N x 12]
//   boxes                           : n_boxes x 3 int32 lattice coordinates, id order
//   occ, vox2id (sparse sets only)  : bitmap and id table over the (gs-1)^3 voxel lattice
#include "../../include/lrf.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include <cuda.h>   // driver API TYPES only (multicast objects, virtual memory management); entry points come from cudaGetDriverEntryPoint

#include "lrf_kernels.cuh"

using namespace lrf;

struct lrf_ctx
{
  int device = 0;
  int gs = 0;
  size_t N = 0;
  float *params = nullptr, *grad = nullptr, *V = nullptr, *S = nullptr;
  float *tv_tmp = nullptr;            // 2*N floats
  int32_t *boxes = nullptr;           // device, n_boxes*3 (null while dense)
  uint32_t *occ = nullptr, *vox2id = nullptr, *occ8 = nullptr;   // occ8: one bit per 8^3-voxel brick, followed by the brick distance field
  uint32_t occ8_words = 0;
  size_t n_boxes = 0;
  bool dense = true;
  float bb_min[3] = {0, 0, 0}, bb_max[3] = {1, 1, 1};
  float view_inv[16], proj_inv[16];
  float lr = 0.15f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8;   // example_tracer.h:120-123
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaStream_t side = nullptr;        // second stream: the TV gradient beside the ray backward
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, tm0 = nullptr, tm1 = nullptr;
  float t_march = 0, t_grad = 0, t_tv = 0, t_adam = 0;
  uint64_t launches = 0;
  // staging
  void *stage = nullptr;
  size_t stage_bytes = 0;
  // fallback queue of the forward render (rays that need the generic march): [0] = count, [1..] = pixel indices
  uint32_t *fb = nullptr;
  size_t fb_cap = 0;
  float *views_dev = nullptr;         // view matrices of a batch render (lrf_render_views_dev_async), n x 16 floats
  size_t views_cap = 0;
  void *hit_scratch = nullptr;        // per-ray hit lists of the packed-batch backward
  size_t hit_bytes = 0;
  int generic_image_grad = 0;   // 1: whole-image backward by the generic thread-per-ray kernel (tests compare the two)
  int force_generic = 0;   // parity taps: 1 = generic march for every ray (tests compare both paths)
  // resident target images: n_targets x tw x th float4, stbi_loadf layout
  float *targets = nullptr;
  float *gamma_lut = nullptr;         // device copy of the 256-entry ldr->hdr table
  uint32_t n_targets = 0, tw = 0, th = 0;
  // multi-GPU exchange over peer memory: parameter / gradient planes of every rank (own rank: the local arrays)
  int peer_world = 0, peer_rank = 0;
  float *peer_param[8] = {nullptr}, *peer_grad[8] = {nullptr};
  bool peer_ipc = false;              // pointers of the other ranks came from cudaIpcOpenMemHandle (must be closed)
  uint32_t *flags = nullptr;          // own barrier flag array: [0..8) one slot per rank, [16] = time-out marker
  uint32_t *peer_flags[8] = {nullptr};
  uint32_t peer_epoch = 0;
  // the exchange through NVLink multicast objects (lrf_mc_*): while attached, params / grad live in cuMemCreate allocations bound to
  // the group's multicast object; mc_va maps it: [0, mc_bytes) parameters of every replica, [mc_bytes, 2 mc_bytes) gradients
  bool mc = false;
  CUmemGenericAllocationHandle mc_mem[2] = {0, 0};   // own physical allocations: params, grad
  CUmemGenericAllocationHandle mc_obj = 0;           // the multicast object
  bool mc_obj_owner = false;
  CUdeviceptr mc_va = 0;
  size_t mc_bytes = 0;                               // padded size of one plane set
  int mc_state = 0;                                  // 0 none, 1 prepared (VMM), 2 object known, 3 device added, 4 bound + mapped
  int mc_unroll = 4, mc_blocks_per_sm = 0;           // options "mc_unroll" (2 / 4 / 8 float4 in flight per thread), "mc_blocks_per_sm" (0 = one block per chunk, else a persistent grid)
  // staged forward render (lrf_staged.cuh): baked 64-byte node records + what was learnt while baking
  float4 *rec = nullptr;              // N records
  size_t rec_N = 0;
  bool rec_dirty = true;              // parameters changed since the last bake
  float *dmax_dev = nullptr;          // largest |density| of the baked grid (device), and its host copy
  float dmax = 0.0f;
  uint32_t *st_stats = nullptr;       // 4 counters, see StagedArgs::stats
  int wray_parallel = 1;              // option "wray_parallel": packed-batch backward enumerates the hits of a group on all lanes at once (lrf_wray.cuh); 0 = lock-step walk
  int lean = 1;                       // option "lean_render": dense colour frames by the register-window march (lrf_lean.cuh); 0 = FastMarch gather kernel
  int staged = 0;                     // option "staged_render": 1 = dense colour renders by the TMA-staged kernel (lrf_staged.cuh), 0 = by the gather kernel
  int tv_blocks_per_sm = 2;           // option "tv_blocks_per_sm": resident TV blocks per SM while it overlaps the colour Adam pass (0 = no cap)
  int percell_overlap = 1;            // option "percell_overlap": lrf_train_percell_async runs {TV -> Adam(density)} beside Adam(colour) on two streams
  int tv_sweep = 1;                   // option "tv_sweep": TV gradient by the TMA z-sweep kernel (lrf_tv.cuh) where it applies; 0 = the 8^3-tile kernel
  int baked = 0;                      // option "baked_records": 1 = the gather kernel reads the 64-byte node records instead of the planes (measured slower: 2.48 vs 2.30 ms per view)
  int staged_nw = 4;                  // option "staged_warps": consumer warps per CTA (4: 16x8 pixel tiles, 8: 16x16)
  int staged_box_w = 0, staged_box_h = 0;   // options "staged_box_w/h": override the footprint box (0 = from the camera)
  int last_render_staged = 0;
  uint32_t *diag_host = nullptr, *diag_dev = nullptr;   // watchdog record of the staged kernel (host-mapped pinned memory)
};

static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return 1; }
#define CK(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                   \
  } while (0)
#define NEED(h) do { if (!(h)) return fail("null context"); CK(cudaSetDevice((h)->device)); } while (0)
#define NEED_GRID(h) do { NEED(h); if (!(h)->params) return fail("grid not initialised (call lrf_init_grid)"); } while (0)

static int ensure_stage(lrf_ctx *h, size_t bytes)
{
  if (h->stage_bytes >= bytes) return 0;
  if (h->stage) cudaFree(h->stage);
  h->stage = nullptr; h->stage_bytes = 0;
  CK(cudaMalloc(&h->stage, bytes));
  h->stage_bytes = bytes;
  return 0;
}

constexpr size_t kFbHeader = 4;   // [0] = fallback count
static int ensure_fb(lrf_ctx *h, size_t rays)
{
  if (h->fb_cap >= rays) return 0;
  if (h->fb) cudaFree(h->fb);
  h->fb = nullptr; h->fb_cap = 0;
  CK(cudaMalloc(&h->fb, (rays + kFbHeader) * sizeof(uint32_t)));
  h->fb_cap = rays;
  return 0;
}

static SceneDev scene_of(const lrf_ctx *h)
{
  SceneDev S{};
  S.gs = h->gs;
  S.n = h->gs - 1;
  S.dens = h->params;
  S.col = reinterpret_cast<const float4 *>(h->params + h->N);
  S.gdens = h->grad;
  S.gcol = h->grad ? reinterpret_cast<float4 *>(h->grad + h->N) : nullptr;
  S.occ = h->dense ? nullptr : h->occ;
  S.vox2id = h->dense ? nullptr : h->vox2id;
  S.occ8 = h->dense ? nullptr : h->occ8;
  S.occ8_words = h->occ8_words;
  S.nb8 = (h->gs - 1 + 7) / 8;
  S.gs_f = (float)h->gs;
  S.inv_gs_f = 1.0f / (float)h->gs;
  S.inv_gs = 1.0 / (double)h->gs;
  const float step_f = (h->bb_max[0] - h->bb_min[0]) / (float)h->gs;   // example_tracer.cpp:649 (float), :692 (same value widened)
  S.step = (double)step_f;
  const double gate = std::sqrt(3.0) * S.step;                          // :145
  float gf = (float)gate;
  if ((double)gf > gate) gf = std::nextafterf(gf, -INFINITY);           // largest float <= gate:  a > gate  <=>  a > gf  for float a
  S.step_gate_f = gf;
  std::memcpy(S.view_inv, h->view_inv, 64);
  std::memcpy(S.proj_inv, h->proj_inv, 64);
  return S;
}


// ---- staged forward render: records, footprint box, tensor maps (device side: lrf_staged.cuh) ----
#ifndef LRF_ST_SWZ
#define LRF_ST_SWZ 0
#endif
typedef CUresult (*lrf_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static lrf_encode_tiled_fn encode_tiled()
{
  // taken from the driver at run time: the library keeps no link-time dependency on libcuda (it must load on a box without a GPU)
  static lrf_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<lrf_encode_tiled_fn>(p);
    cudaGetLastError();
  }
  return fn;
}

static void params_changed(lrf_ctx *h) { h->rec_dirty = true; }
static int fail_with_diag(lrf_ctx *h, const char *what_call, cudaError_t e)
{
  std::string m = std::string(what_call) + ": " + cudaGetErrorString(e);
  if (h->diag_host && h->diag_host[0]) {
    const uint32_t *d = h->diag_host;
    const char *what = d[0] == 1 ? "producer: ring never freed (q, qmin, p0)" : d[0] == 2 ? "producer: drain (plane, issued, parity)"
                       : d[0] == 3 ? "consumer: plane never arrived (qready, qmin, parity)" : d[0] == 4 ? "producer: own TMA never completed (plane, q, parity)" : d[0] == 5 ? "dynamic shared memory is not 1024-byte aligned (address)" : d[0] == 6 ? "tv_sweep_kernel: density plane never arrived (step, block, parity)" : "?";
    m += " | staged-render watchdog: " + std::string(what) + " block " + std::to_string(d[1]) + " thread " + std::to_string(d[2]) + " values " +
         std::to_string((int)d[3]) + " " + std::to_string((int)d[4]) + " " + std::to_string((int)d[5]) + " active-mask " + std::to_string(d[6]);
    m += " | dump:";
    for (int i = 8; i < 96; i++) m += " " + std::to_string((int)d[i]);
  }
  return fail(m);
}
#define CKD(h, call)                                                  \
  do {                                                                \
    cudaError_t e_ = (call);                                          \
    if (e_ != cudaSuccess) return fail_with_diag(h, #call, e_);       \
  } while (0)
static int sync_with_diag(lrf_ctx *h)
{
  CKD(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// ray of a pixel centre, the arithmetic of gen_ray (lrf_device.cuh) — only used to size the footprint box, rounding is immaterial here
static void host_gen_ray(const float *view_inv, const float *proj_inv, double x, double y, uint32_t w, uint32_t hgt, double *o, double *d)
{
  auto mv = [](double *res, const float *B, const double *V) {
    for (int r = 0; r < 4; r++) res[r] = V[0] * B[r] + V[1] * B[4 + r] + V[2] * B[8 + r] + V[3] * B[12 + r];
  };
  const double u = (x + 0.5) / w, v = (y + 0.5) / hgt;
  double pos[4] = {2 * u - 1, 2 * v - 1, 0, 1}, p2[4];
  mv(p2, proj_inv, pos);
  for (int i = 0; i < 3; i++) p2[i] /= p2[3];
  double len = std::sqrt(p2[0] * p2[0] + p2[1] * p2[1] + p2[2] * p2[2]);
  double rd[4] = {p2[0] / len, p2[1] / len, p2[2] / len, 0}, rp[4] = {0, 0, 0, 1}, rpt[4], rdt[4];
  mv(rpt, view_inv, rp);
  mv(rdt, view_inv, rd);
  len = std::sqrt(rdt[0] * rdt[0] + rdt[1] * rdt[1] + rdt[2] * rdt[2] + rdt[3] * rdt[3]);
  for (int i = 0; i < 3; i++) { o[i] = rpt[i]; d[i] = rdt[i] / len; }
}

// Footprint box (nodes) a tile of tw x th pixels needs on a node plane: the same corner-ray construction as the producer warp of
// render_staged_kernel, evaluated on a 13 x 13 sample of tiles.  A tile that needs more than this sends its pixels to the exact kernel,
// so the estimate only has to be good, not safe.
static void staged_box_need(const lrf_ctx *h, uint32_t w, uint32_t hgt, int tw, int th, int Wn[3], int Hn[3])
{
  for (int a = 0; a < 3; a++) { Wn[a] = 0; Hn[a] = 0; }
  const int gs = h->gs;
  const uint32_t ntx = (w + tw - 1) / tw, nty = (hgt + th - 1) / th;
  const int NS = 12;
  for (int iy = 0; iy <= NS; iy++)
    for (int ix = 0; ix <= NS; ix++) {
      const uint32_t tx = (uint32_t)((uint64_t)(ntx - 1) * ix / NS), ty = (uint32_t)((uint64_t)(nty - 1) * iy / NS);
      const double x0 = (double)tx * tw, y0 = (double)ty * th;
      double oc[3], dc[3];
      host_gen_ray(h->view_inv, h->proj_inv, std::min<double>(x0 + tw / 2, w - 1), std::min<double>(y0 + th / 2, hgt - 1), w, hgt, oc, dc);
      const double ax = std::fabs(dc[0]), ay = std::fabs(dc[1]), az = std::fabs(dc[2]);
      const int A = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2), U = A == 0 ? 1 : 0, V = A == 2 ? 1 : 2;
      double lo_u[2] = {1e30, 1e30}, hi_u[2] = {-1e30, -1e30}, lo_v[2] = {1e30, 1e30}, hi_v[2] = {-1e30, -1e30};
      bool ok = true;
      for (int c = 0; c < 4; c++) {
        double o[3], d[3];
        host_gen_ray(h->view_inv, h->proj_inv, std::min<double>(x0 + ((c & 1) ? tw - 1 : 0), w - 1), std::min<double>(y0 + ((c & 2) ? th - 1 : 0), hgt - 1), w,
                     hgt, o, d);
        if (!(std::fabs(d[A]) >= 0.25) || (d[A] > 0) != (dc[A] > 0)) { ok = false; break; }
        const double su = d[U] / d[A], sv = d[V] / d[A], cu = o[U] * gs - o[A] * gs * su, cv = o[V] * gs - o[A] * gs * sv;
        for (int e = 0; e < 2; e++) {
          const double k = e ? gs - 1 : 0, fu = cu + k * su, fv = cv + k * sv;
          lo_u[e] = std::min(lo_u[e], std::floor(fu - std::fabs(su) - 0.03125)); hi_u[e] = std::max(hi_u[e], std::floor(fu + std::fabs(su) + 0.03125) + 1);
          lo_v[e] = std::min(lo_v[e], std::floor(fv - std::fabs(sv) - 0.03125)); hi_v[e] = std::max(hi_v[e], std::floor(fv + std::fabs(sv) + 0.03125) + 1);
        }
      }
      if (!ok) continue;
      for (int e = 0; e < 2; e++) {
        Wn[A] = std::max(Wn[A], (int)std::min(1e6, hi_u[e] - lo_u[e] + 1));
        Hn[A] = std::max(Hn[A], (int)std::min(1e6, hi_v[e] - lo_v[e] + 1));
      }
    }
}

// TV gradient of the nodes with z in [z0, z1) on stream st: the TMA z-sweep (lrf_tv.cuh) where it applies, else the 8^3-tile kernel
static int ensure_diag(lrf_ctx *h)
{
  if (h->diag_host) return 0;
  CK(cudaHostAlloc((void **)&h->diag_host, 128 * sizeof(uint32_t), cudaHostAllocMapped));
  std::memset(h->diag_host, 0, 128 * sizeof(uint32_t));
  CK(cudaHostGetDevicePointer((void **)&h->diag_dev, h->diag_host, 0));
  return 0;
}
static bool tv_sweep_map(lrf_ctx *h, CUtensorMap &tm)
{
  if (!h->tv_sweep || h->gs < 64 || (h->gs % 32) != 0) return false;
  lrf_encode_tiled_fn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t gs = (cuuint64_t)h->gs;
  const cuuint64_t dims[3] = {gs, gs, gs}, strides[2] = {4 * gs, 4 * gs * gs};
  const cuuint32_t box[3] = {40, 12, 1}, estr[3] = {1, 1, 1};   // kSwDX x kSwDY of lrf_tv.cuh
  return enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, h->params, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static void tv_dispatch(lrf_ctx *h, int z0, int z1, int atomic, int max_blocks, bool whole, cudaStream_t st, int blocks_per_sm = 0)
{
  CUtensorMap tm;
  if (tv_sweep_map(h, tm) && ensure_diag(h) == 0) {
    launch_tv_sweep(tm, h->gs, h->grad, z0, z1, atomic, h->diag_dev, st, blocks_per_sm);
    return;
  }
  if (whole) launch_tv_grad(h->gs, h->params, h->grad, h->tv_tmp, h->tv_tmp + h->N, st);
  else launch_tv_grad_slab(h->gs, h->params, h->grad, z0, z1, atomic, max_blocks, st);
}

// the baked 64-byte node records both forward kernels can read (lrf_staged.cuh), rebuilt when the parameters changed.
// 0 = ready, 1 = error, 2 = no room for the second copy of the grid (the caller renders from the planes)
static int ensure_records(lrf_ctx *h)
{
  if (!h->rec || h->rec_N != h->N) {
    cudaFree(h->rec);
    h->rec = nullptr; h->rec_N = 0;
    if (cudaMalloc(&h->rec, h->N * 64) != cudaSuccess) { cudaGetLastError(); return 2; }
    h->rec_N = h->N;
    h->rec_dirty = true;
  }
  if (!h->dmax_dev) CK(cudaMalloc(&h->dmax_dev, sizeof(float)));
  if (h->rec_dirty) {
    launch_bake_records(scene_of(h), h->rec, h->dmax_dev, h->stream);
    h->launches++;
    CK(cudaMemcpyAsync(&h->dmax, h->dmax_dev, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->rec_dirty = false;
  }
  return 0;
}

// 0 = T is ready and the staged kernel can run; 1 = error (g_err set); 2 = this render is not eligible (caller uses the gather kernel)
static int staged_prepare(lrf_ctx *h, uint32_t w, uint32_t hgt, StagedArgs &T)
{
  if (!h->dense || h->gs < 8 || h->gs > 512) return 2;   // 10-bit step counters; tiny grids gain nothing
  lrf_encode_tiled_fn enc = encode_tiled();
  if (!enc) return 2;
  const int nw = h->staged_nw == 8 ? 8 : 4;
  int Wn[3], Hn[3];
  staged_box_need(h, w, hgt, 16, 2 * nw, Wn, Hn);
  uint32_t slot = 0;
  for (int A = 0; A < 3; A++) {
    // an axis no sampled tile marches along still gets a (small) box: a tile that does use it and does not fit goes to the exact kernel
    int W = h->staged_box_w > 0 ? h->staged_box_w : std::max(4, Wn[A] + 1), H = h->staged_box_h > 0 ? h->staged_box_h : std::max(4, Hn[A] + 1);
    if (h->staged_box_w <= 0)
      while ((W & 7) == 0 || (W & 7) == 1 || (W & 7) == 7 || (W & 7) == 4) W++;   // row pitch in records: neighbours within +-2 rows / columns land in different bank groups
    if (std::getenv("LRF_DEBUG")) std::fprintf(stderr, "[lrf] staged box A=%d: need %d x %d -> W %d H %d (nw %d, %ux%u)\n", A, Wn[A], Hn[A], W, H, nw, w, hgt);
    if (W > 64 || H > 64) return 2;
    T.W[A] = W; T.H[A] = H;
    T.plane_bytes[A] = (uint32_t)W * H * 64u;
    slot = std::max(slot, (T.plane_bytes[A] + 511u) & ~511u);
  }
  if (1024 + 512 + 8 * (size_t)slot > 200 * 1024) return 2;
  { const int r = ensure_records(h); if (r) return r; }
  if (!h->st_stats) { CK(cudaMalloc(&h->st_stats, 4 * sizeof(uint32_t))); CK(cudaMemsetAsync(h->st_stats, 0, 4 * sizeof(uint32_t), h->stream)); }
  T.slot_bytes = slot;
  T.err_sigma = 16.0f * 0x1p-24f * h->dmax;   // NaN / inf densities make every break ambiguous: the exact kernel then renders those rays
  T.stats = h->st_stats;
  if (ensure_diag(h)) return 1;
  T.diag = h->diag_dev;
  const cuuint64_t gs = (cuuint64_t)h->gs;
  const cuuint64_t dims[4] = {16, gs, gs, gs}, strides[3] = {64, 64 * gs, 64 * gs * gs};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int A = 0; A < 3; A++) {
    const cuuint32_t W = (cuuint32_t)T.W[A], H = (cuuint32_t)T.H[A];
    const cuuint32_t box[4] = {16, A == 0 ? 1u : W, A == 0 ? W : (A == 1 ? 1u : H), A == 2 ? 1u : H};
    const CUresult r = enc(&T.tm[A], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, h->rec, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           LRF_ST_SWZ ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed: CUresult " + std::to_string((int)r));
  }
  return 0;
}
static int render_dispatch(lrf_ctx *h, uint32_t *out_dev, uint32_t w, uint32_t hgt, int grey, uint32_t *tap_shaded)
{
  h->last_render_staged = 0;
  if (h->staged && !grey) {
    StagedArgs T;
    const int r = staged_prepare(h, w, hgt, T);
    if (r == 1) return 1;
    if (r == 0) {
      const int e = launch_render_staged(scene_of(h), T, h->staged_nw == 8 ? 8 : 4, out_dev, w, hgt, h->fb, h->fb + kFbHeader, tap_shaded, h->stream);
      if (e) return fail(std::string("staged render: ") + cudaGetErrorString((cudaError_t)e));
      h->launches += 2;   // render_staged_kernel + render_list_kernel
      h->last_render_staged = 1;
      return 0;
    }
  }
  if (tap_shaded) return fail("the staged render is not available for this configuration (dense grids of 8..512 nodes per axis, colour mode)");
  SceneDev S = scene_of(h);
  if (h->baked && h->dense && !grey) {   // gather kernel over the baked records (fewer L1 wavefronts per node); from the planes if they do not fit
    const int r = ensure_records(h);
    if (r == 1) return 1;
    if (r == 0) S.rec = h->rec;
  }
  launch_render_image(S, out_dev, w, hgt, grey, h->fb, h->fb + kFbHeader, h->stream, 0, 0xffffffffu, h->lean);
  h->launches += 2;   // render_lean_kernel / render_image_kernel + render_fallback_kernel
  return 0;
}

static void mc_teardown(lrf_ctx *h);
static int peer_check(lrf_ctx *h, int world, int rank)
{
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail("peer group: world must be 1..8 and 0 <= rank < world");
  if (h->gs % (8 * world)) return fail("peer group: gridSize must be a multiple of 8 * world (z-slabs of whole 8^3 TV tiles)");
  return 0;
}
static void peer_detach(lrf_ctx *h)
{
  if (h->peer_ipc)
    for (int r = 0; r < h->peer_world; r++)
      if (r != h->peer_rank) {
        if (h->peer_param[r]) cudaIpcCloseMemHandle(h->peer_param[r]);
        if (h->peer_grad[r]) cudaIpcCloseMemHandle(h->peer_grad[r]);
        if (h->peer_flags[r]) cudaIpcCloseMemHandle(h->peer_flags[r]);
      }
  mc_teardown(h);
  for (int r = 0; r < 8; r++) { h->peer_flags[r] = nullptr; h->peer_param[r] = nullptr; h->peer_grad[r] = nullptr; }
  h->peer_world = 0;
  h->peer_ipc = false;
}

static void free_grid(lrf_ctx *h)
{
  peer_detach(h);   // the attached pointers describe replicas of the grid that is going away
  cudaFree(h->params); cudaFree(h->grad); cudaFree(h->V); cudaFree(h->S); cudaFree(h->tv_tmp);
  cudaFree(h->boxes); cudaFree(h->occ); cudaFree(h->vox2id); cudaFree(h->occ8);
  cudaFree(h->rec);
  h->rec = nullptr; h->rec_N = 0; h->rec_dirty = true;
  h->params = h->grad = h->V = h->S = h->tv_tmp = nullptr;
  h->boxes = nullptr; h->occ = h->vox2id = h->occ8 = nullptr;
}

static int alloc_grid(lrf_ctx *h, int gs)
{
  free_grid(h);
  h->gs = gs;
  h->N = (size_t)gs * gs * gs;
  const size_t bytes = h->N * kCell * sizeof(float);
  cudaError_t e = cudaMalloc(&h->params, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->grad, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->V, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->S, bytes);
  if (e != cudaSuccess) {   // e.g. 1024^3: 4 x 56 GB does not fit — leave the context without a grid rather than half-built
    cudaGetLastError();
    free_grid(h);
    h->gs = 0;
    h->N = 0;
    h->n_boxes = 0;
    return fail(std::string("cannot allocate a ") + std::to_string(gs) + "^3 grid (4 arrays of " + std::to_string(bytes >> 20) + " MB): " +
                cudaGetErrorString(e));
  }
  CK(cudaMemsetAsync(h->grad, 0, bytes, h->stream));
  CK(cudaMemsetAsync(h->V, 0, bytes, h->stream));
  CK(cudaMemsetAsync(h->S, 0, bytes, h->stream));
  h->dense = true;
  h->n_boxes = (size_t)(gs - 1) * (gs - 1) * (gs - 1);
  h->rec_dirty = true;
  return 0;
}


// ------------------------------------------------------------------------------------------------
// NVLink multicast ("NVLS") plumbing for the training exchange (adam_mc_kernel, lrf_kernels.cu).  No reference counterpart (the
// reference is one process, #pragma omp parallel for at example_tracer.cpp:668).  Driver entry points are taken at run time.
namespace {
struct McDriver
{
  CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
  CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
  bool ok = false;
};
template <class F> bool mc_sym(const char *name, F &fn)
{
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}
McDriver &mc_driver()
{
  static McDriver D;
  static bool tried = false;
  if (!tried) {
    tried = true;
    D.ok = mc_sym("cuMemCreate", D.MemCreate) && mc_sym("cuMemRelease", D.MemRelease) && mc_sym("cuMemAddressReserve", D.MemAddressReserve) &&
           mc_sym("cuMemAddressFree", D.MemAddressFree) && mc_sym("cuMemMap", D.MemMap) && mc_sym("cuMemUnmap", D.MemUnmap) &&
           mc_sym("cuMemSetAccess", D.MemSetAccess) && mc_sym("cuMemGetAllocationGranularity", D.MemGetAllocationGranularity) &&
           mc_sym("cuMemExportToShareableHandle", D.MemExportToShareableHandle) &&
           mc_sym("cuMemImportFromShareableHandle", D.MemImportFromShareableHandle) && mc_sym("cuMulticastCreate", D.MulticastCreate) &&
           mc_sym("cuMulticastAddDevice", D.MulticastAddDevice) && mc_sym("cuMulticastBindMem", D.MulticastBindMem) &&
           mc_sym("cuMulticastUnbind", D.MulticastUnbind) && mc_sym("cuMulticastGetGranularity", D.MulticastGetGranularity) &&
           mc_sym("cuDeviceGet", D.DeviceGet) && mc_sym("cuDeviceGetAttribute", D.DeviceGetAttribute) && mc_sym("cuGetErrorString", D.GetErrorString);
  }
  return D;
}
int mc_fail(const char *what, CUresult r)
{
  const char *s = nullptr;
  if (mc_driver().GetErrorString) mc_driver().GetErrorString(r, &s);
  return fail(std::string(what) + ": CUresult " + std::to_string((int)r) + (s ? std::string(" (") + s + ")" : std::string()));
}
#define CKU(call)                                              \
  do {                                                         \
    CUresult r_ = (call);                                      \
    if (r_ != CUDA_SUCCESS) return mc_fail(#call, r_);         \
  } while (0)
CUmemAllocationProp mc_alloc_prop(int device)
{
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}
// a plane set (13 N floats) in a cuMemCreate allocation of `bytes`, mapped read-write on the context's device
int mc_alloc_planes(lrf_ctx *h, size_t bytes, size_t gran, CUmemGenericAllocationHandle &mem, float *&ptr)
{
  McDriver &D = mc_driver();
  const CUmemAllocationProp prop = mc_alloc_prop(h->device);
  CKU(D.MemCreate(&mem, bytes, &prop, 0));
  CUdeviceptr va = 0;
  CKU(D.MemAddressReserve(&va, bytes, gran, 0, 0));
  CKU(D.MemMap(va, bytes, 0, mem, 0));
  CUmemAccessDesc acc{};
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CKU(D.MemSetAccess(va, bytes, &acc, 1));
  ptr = reinterpret_cast<float *>(va);
  return 0;
}
void mc_free_planes(size_t bytes, CUmemGenericAllocationHandle mem, float *ptr)
{
  McDriver &D = mc_driver();
  if (ptr) { D.MemUnmap((CUdeviceptr)ptr, bytes); D.MemAddressFree((CUdeviceptr)ptr, bytes); }
  if (mem) D.MemRelease(mem);
}
}  // namespace

// leaves the group: the multicast mapping goes away and the planes move back into cudaMalloc allocations (contents kept)
static void mc_teardown(lrf_ctx *h)
{
  if (!h->mc) return;
  McDriver &D = mc_driver();
  cudaSetDevice(h->device);
  const bool dbg = std::getenv("LRF_DEBUG") != nullptr;
  auto step = [&](const char *what, long rc) {
    const cudaError_t e = cudaDeviceSynchronize();
    if (dbg || e != cudaSuccess || rc != 0) std::fprintf(stderr, "[lrf] mc_teardown dev %d: %s -> rc %ld, device %s\n", h->device, what, rc, cudaGetErrorString(e));
  };
  step("entry", 0);
  // 1. the contents move to ordinary allocations while everything is still mapped
  const size_t bytes = h->N * kCell * sizeof(float);
  float *np = nullptr, *ng = nullptr;
  if (h->params) {
    long rc = (long)cudaMalloc(&np, bytes);
    if (!rc) rc = (long)cudaMalloc(&ng, bytes);
    if (!rc) rc = (long)cudaMemcpy(np, h->params, bytes, cudaMemcpyDeviceToDevice);
    if (!rc) rc = (long)cudaMemcpy(ng, h->grad, bytes, cudaMemcpyDeviceToDevice);
    step("copy out", rc);
  }
  // 2. the multicast mapping, the bindings, the object
  if (h->mc_va) {
    long rc = (long)D.MemUnmap(h->mc_va, 2 * h->mc_bytes);
    if (!rc) rc = (long)D.MemAddressFree(h->mc_va, 2 * h->mc_bytes);
    step("unmap multicast range", rc);
  }
  if (h->mc_obj) {
    if (h->mc_state >= 4) {
      CUdevice dev;
      long rc = (long)D.DeviceGet(&dev, h->device);
      if (!rc) rc = (long)D.MulticastUnbind(h->mc_obj, dev, 0, h->mc_bytes);
      if (!rc) rc = (long)D.MulticastUnbind(h->mc_obj, dev, h->mc_bytes, h->mc_bytes);
      step("unbind", rc);
    }
    step("release object", (long)D.MemRelease(h->mc_obj));
  }
  // 3. the cuMemCreate planes
  mc_free_planes(h->mc_bytes, h->mc_mem[0], h->params);
  mc_free_planes(h->mc_bytes, h->mc_mem[1], h->grad);
  step("free planes", 0);
  h->params = np; h->grad = ng;
  h->mc = false; h->mc_obj = 0; h->mc_va = 0; h->mc_mem[0] = h->mc_mem[1] = 0; h->mc_bytes = 0; h->mc_state = 0; h->mc_obj_owner = false;
  params_changed(h);
}

extern "C" {
int lrf_mc_supported(int device)
{
  McDriver &D = mc_driver();
  if (!D.ok) return 0;
  CUdevice dev;
  int v = 0;
  if (D.DeviceGet(&dev, device) != CUDA_SUCCESS) return 0;
  if (D.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return 0;
  return v;
}
// step 1 (every rank): the parameter and gradient planes move into shareable cuMemCreate allocations padded to the multicast granularity
int lrf_mc_prepare(lrf_ctx *h, int world, int rank)
{
  NEED_GRID(h);
  if (peer_check(h, world, rank)) return 1;
  if (!lrf_mc_supported(h->device)) return fail("NVLink multicast objects are not supported on this device / driver");
  peer_detach(h);
  McDriver &D = mc_driver();
  CUmulticastObjectProp mp{};
  mp.numDevices = (unsigned)world;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  const size_t bytes = h->N * kCell * sizeof(float);
  mp.size = bytes;
  size_t gran = 0, agran = 0;
  CKU(D.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  const CUmemAllocationProp prop = mc_alloc_prop(h->device);
  CKU(D.MemGetAllocationGranularity(&agran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (agran > gran) gran = agran;
  const size_t padded = (bytes + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle m0 = 0, m1 = 0;
  float *np = nullptr, *ng = nullptr;
  if (mc_alloc_planes(h, padded, gran, m0, np)) return 1;
  if (mc_alloc_planes(h, padded, gran, m1, ng)) { mc_free_planes(padded, m0, np); return 1; }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(np, h->params, bytes, cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(ng, h->grad, bytes, cudaMemcpyDeviceToDevice));
  cudaFree(h->params); cudaFree(h->grad);
  h->params = np; h->grad = ng;
  h->mc_mem[0] = m0; h->mc_mem[1] = m1; h->mc_bytes = padded;
  h->mc = true; h->mc_state = 1;
  h->peer_world = world; h->peer_rank = rank; h->peer_ipc = true;   // the flag barrier maps its arrays with CUDA IPC
  params_changed(h);
  return 0;
}
// step 2, ONE rank: create the group's multicast object and export it as a POSIX file descriptor (to be passed to the other ranks'
// processes, e.g. over a unix socket with SCM_RIGHTS); the caller closes the descriptor
int lrf_mc_create(lrf_ctx *h, int *fd_out)
{
  NEED_GRID(h);
  if (!h->mc || h->mc_state != 1) return fail("lrf_mc_create: call lrf_mc_prepare first");
  McDriver &D = mc_driver();
  CUmulticastObjectProp mp{};
  mp.numDevices = (unsigned)h->peer_world;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  mp.size = 2 * h->mc_bytes;
  CKU(D.MulticastCreate(&h->mc_obj, &mp));
  h->mc_obj_owner = true;
  h->mc_state = 2;
  if (fd_out) {
    int fd = -1;
    CKU(D.MemExportToShareableHandle(&fd, h->mc_obj, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *fd_out = fd;
  }
  return 0;
}
// step 2, the other ranks: import the object from the descriptor received
int lrf_mc_import(lrf_ctx *h, int fd)
{
  NEED_GRID(h);
  if (!h->mc || h->mc_state != 1) return fail("lrf_mc_import: call lrf_mc_prepare first");
  McDriver &D = mc_driver();
  CKU(D.MemImportFromShareableHandle(&h->mc_obj, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  h->mc_state = 2;
  return 0;
}
// step 3 (every rank; all ranks must have done it before anyone binds): add the own device to the object
int lrf_mc_add_device(lrf_ctx *h)
{
  NEED_GRID(h);
  if (!h->mc || h->mc_state != 2) return fail("lrf_mc_add_device: no multicast object yet (lrf_mc_create / lrf_mc_import)");
  McDriver &D = mc_driver();
  CUdevice dev;
  CKU(D.DeviceGet(&dev, h->device));
  CKU(D.MulticastAddDevice(h->mc_obj, dev));
  h->mc_state = 3;
  return 0;
}
// step 4 (every rank, after a group-wide barrier): bind the own planes and map the multicast range
int lrf_mc_bind(lrf_ctx *h)
{
  NEED_GRID(h);
  if (!h->mc || h->mc_state != 3) return fail("lrf_mc_bind: add the device first (lrf_mc_add_device)");
  McDriver &D = mc_driver();
  CKU(D.MulticastBindMem(h->mc_obj, 0, h->mc_mem[0], 0, h->mc_bytes, 0));
  CKU(D.MulticastBindMem(h->mc_obj, h->mc_bytes, h->mc_mem[1], 0, h->mc_bytes, 0));
  CUmulticastObjectProp mp{};
  mp.numDevices = (unsigned)h->peer_world;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  mp.size = 2 * h->mc_bytes;
  size_t gran = 0;
  CKU(D.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  CKU(D.MemAddressReserve(&h->mc_va, 2 * h->mc_bytes, gran, 0, 0));
  CKU(D.MemMap(h->mc_va, 2 * h->mc_bytes, 0, h->mc_obj, 0));
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = h->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CKU(D.MemSetAccess(h->mc_va, 2 * h->mc_bytes, &acc, 1));
  h->mc_state = 4;
  return 0;
}
// all four steps for a group whose contexts live in ONE process (tests, single-process multi-GPU hosts); flags as lrf_peer_attach_local
int lrf_mc_attach_local(int world, lrf_ctx *const *ctxs)
{
  if (world < 1 || world > 8 || !ctxs) return fail("lrf_mc_attach_local: world must be 1..8");
  for (int r = 0; r < world; r++) {
    if (!ctxs[r]) return fail("null context");
    if (lrf_mc_prepare(ctxs[r], world, r)) return 1;
    ctxs[r]->peer_ipc = false;
  }
  if (lrf_mc_create(ctxs[0], nullptr)) return 1;
  for (int r = 1; r < world; r++) {
    // one process: the handle itself is valid for every device; every context keeps (and later releases) its own reference
    int fd = -1;
    McDriver &D = mc_driver();
    CKU(D.MemExportToShareableHandle(&fd, ctxs[0]->mc_obj, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    const int rc = lrf_mc_import(ctxs[r], fd);
    close(fd);
    if (rc) return 1;
  }
  for (int r = 0; r < world; r++)
    if (lrf_mc_add_device(ctxs[r])) return 1;
  for (int r = 0; r < world; r++)
    if (lrf_mc_bind(ctxs[r])) return 1;
  for (int r = 0; r < world; r++)
    for (int q = 0; q < world; q++) ctxs[r]->peer_flags[q] = ctxs[q]->flags;
  return 0;
}
int lrf_mc_attached(lrf_ctx *h) { return h && h->mc && h->mc_state == 4 ? 1 : 0; }
}

extern "C" {

const char *lrf_last_error(void) { return g_err.c_str(); }

int lrf_create(int device, lrf_ctx **out)
{
  if (!out) return fail("null out");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
  if (device < 0 || device >= count) return fail("device index out of range");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + "; this build carries sm_100a code only");
  lrf_ctx *h = new lrf_ctx();
  h->device = device;
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  {
    // the second stream carries the light kernels that run BESIDE an HBM-bound one (the TV gradient next to Adam or next to the ray
    // backward): highest priority, so that its blocks are placed as soon as an SM has room instead of behind the other kernel's queue
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CK(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, greatest));
  }
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CK(cudaMalloc(&h->flags, 32 * sizeof(uint32_t)));
  CK(cudaMemset(h->flags, 0, 32 * sizeof(uint32_t)));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  CK(cudaEventCreate(&h->tm0));
  CK(cudaEventCreate(&h->tm1));
  // defaults of the reference constructor (example_tracer.h:43-49): caller overrides through the setters
  for (int i = 0; i < 16; i++) h->view_inv[i] = h->proj_inv[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  // experiment knobs (tools/variants.py): the same switches as lrf_set_option
  if (const char *e = std::getenv("LRF_STAGED")) h->staged = std::atoi(e) ? 1 : 0;
  if (const char *e = std::getenv("LRF_BAKED")) h->baked = std::atoi(e) ? 1 : 0;
  if (const char *e = std::getenv("LRF_STAGED_WARPS")) h->staged_nw = std::atoi(e) == 8 ? 8 : 4;
  if (const char *e = std::getenv("LRF_STAGED_BOX_W")) h->staged_box_w = std::atoi(e);
  if (const char *e = std::getenv("LRF_STAGED_BOX_H")) h->staged_box_h = std::atoi(e);
  *out = h;
  return 0;
}

int lrf_destroy(lrf_ctx *h)
{
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  free_grid(h);
  cudaFree(h->stage);
  cudaFree(h->fb);
  cudaFree(h->views_dev);
  cudaFree(h->hit_scratch);
  cudaFree(h->targets);
  cudaFree(h->gamma_lut);
  cudaFree(h->dmax_dev);
  cudaFree(h->st_stats);
  if (h->diag_host) cudaFreeHost(h->diag_host);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->tm0); cudaEventDestroy(h->tm1);
  cudaFree(h->flags);
  cudaStreamDestroy(h->side);
  cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int lrf_sync(lrf_ctx *h) { NEED(h); return sync_with_diag(h); }

int lrf_init_grid(lrf_ctx *h, int gs)
{
  NEED(h);
  if (gs < 2 || gs > 1024 || (gs & (gs - 1))) return fail("gridSize must be a power of two in [2, 1024]");
  if (alloc_grid(h, gs)) return 1;
  launch_fill(h->params, h->N, 0.01f, h->stream);               // density  (example_tracer.h:65)
  launch_fill(h->params + h->N, h->N * 12, 0.1f, h->stream);    // RGB + SH (example_tracer.h:66-67)
  h->launches += 2;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// SURVEY 8(d) synthetic grid, generated on the device (bit-identical to the host generator the tests use)
int lrf_fill_synthetic(lrf_ctx *h, float density_scale, int mode)
{
  NEED_GRID(h);
  if (mode < 0 || mode > 2) return fail("mode must be 0 (LCG), 1 (LCG x Gaussian blob) or 2 (sphere: positive inside r = 0.288, negative outside)");
  launch_fill_synthetic(h->gs, density_scale, mode, h->params, h->stream);
  params_changed(h);
  h->launches++;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int lrf_init_grad(lrf_ctx *h) { return lrf_zero_grad(h); }   // resize() value-initialises in the reference

int lrf_zero_grad(lrf_ctx *h)
{
  NEED_GRID(h);
  launch_fill(h->grad, h->N * kCell, 0.0f, h->stream);
  h->launches++;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int lrf_set_bbox(lrf_ctx *h, const float mn[3], const float mx[3])
{
  NEED(h);
  for (int a = 0; a < 3; a++) { h->bb_min[a] = mn[a]; h->bb_max[a] = mx[a]; }
  return 0;
}
int lrf_grid_size(lrf_ctx *h) { return h ? h->gs : 0; }

static int upload13(lrf_ctx *h, float *dst_planes, const float *src_host)
{
  NEED_GRID(h);
  const size_t bytes = h->N * kCell * sizeof(float);
  if (ensure_stage(h, bytes)) return 1;
  CK(cudaMemcpyAsync(h->stage, src_host, bytes, cudaMemcpyHostToDevice, h->stream));
  launch_aos_to_planes((const float *)h->stage, dst_planes, h->N, h->stream);
  h->launches++;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
static int download13(lrf_ctx *h, const float *src_planes, float *dst_host)
{
  NEED_GRID(h);
  const size_t bytes = h->N * kCell * sizeof(float);
  if (ensure_stage(h, bytes)) return 1;
  launch_planes_to_aos(src_planes, (float *)h->stage, h->N, h->stream);
  h->launches++;
  CK(cudaMemcpyAsync(dst_host, h->stage, bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int lrf_grid_upload(lrf_ctx *h, const float *c)
{
  if (!h) return fail("null context");
  params_changed(h);
  return upload13(h, h->params, c);
}
int lrf_grid_download(lrf_ctx *h, float *c) { return h ? download13(h, h->params, c) : fail("null context"); }
int lrf_grad_upload(lrf_ctx *h, const float *c) { return h ? upload13(h, h->grad, c) : fail("null context"); }
int lrf_grad_download(lrf_ctx *h, float *c) { return h ? download13(h, h->grad, c) : fail("null context"); }
int lrf_adam_state_upload(lrf_ctx *h, int which, const float *c) { return h ? upload13(h, which ? h->S : h->V, c) : fail("null context"); }
int lrf_adam_state_download(lrf_ctx *h, int which, float *c) { return h ? download13(h, which ? h->S : h->V, c) : fail("null context"); }

static int install_boxes_dev(lrf_ctx *h, int32_t *xyz_dev, size_t n)  // takes ownership of xyz_dev
{
  cudaFree(h->boxes); cudaFree(h->occ); cudaFree(h->vox2id); cudaFree(h->occ8);
  h->boxes = xyz_dev;
  h->occ = h->vox2id = h->occ8 = nullptr;
  h->n_boxes = n;
  h->dense = false;
  const size_t nv = (size_t)(h->gs - 1) * (h->gs - 1) * (h->gs - 1);
  const size_t words = (nv + 31) / 32 + 1;
  CK(cudaMalloc(&h->occ, words * 4));
  CK(cudaMalloc(&h->vox2id, (nv + 1) * 4));
  CK(cudaMemsetAsync(h->occ, 0, words * 4, h->stream));
  CK(cudaMemsetAsync(h->vox2id, 0, (nv + 1) * 4, h->stream));
  const int nb8 = (h->gs - 1 + 7) / 8;
  const size_t words8 = ((size_t)nb8 * nb8 * nb8 + 31) / 32 + 1;
  const size_t nb3 = (size_t)nb8 * nb8 * nb8;
  CK(cudaMalloc(&h->occ8, words8 * 4 + 2 * nb3 + 16));      // coarse bitmap + the brick distance field (one byte per brick) + its ping-pong copy
  CK(cudaMemsetAsync(h->occ8, 0, words8 * 4, h->stream));
  h->occ8_words = (uint32_t)words8;
  launch_build_box_maps(h->gs - 1, h->boxes, n, h->occ, h->vox2id, h->occ8, nb8, h->stream);
  launch_brick_dist(nb8, h->occ8, reinterpret_cast<uint8_t *>(h->occ8 + words8), reinterpret_cast<uint8_t *>(h->occ8 + words8) + nb3, h->stream);
  h->launches += 2;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int lrf_set_boxes(lrf_ctx *h, const int32_t *xyz, size_t n)
{
  NEED_GRID(h);
  const int nv = h->gs - 1;
  for (size_t i = 0; i < 3 * n; i++)
    if (xyz[i] < 0 || xyz[i] >= nv) return fail("box coordinate outside the (gs-1)^3 lattice");
  int32_t *dev = nullptr;
  CK(cudaMalloc(&dev, (n ? n : 1) * 3 * sizeof(int32_t)));
  if (n) CK(cudaMemcpyAsync(dev, xyz, n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  return install_boxes_dev(h, dev, n);
}
int lrf_set_boxes_dense(lrf_ctx *h)
{
  NEED_GRID(h);
  cudaFree(h->boxes); cudaFree(h->occ); cudaFree(h->vox2id); cudaFree(h->occ8);
  h->boxes = nullptr; h->occ = h->vox2id = h->occ8 = nullptr;
  h->dense = true;
  h->n_boxes = (size_t)(h->gs - 1) * (h->gs - 1) * (h->gs - 1);
  return 0;
}
size_t lrf_box_count(lrf_ctx *h) { return h ? h->n_boxes : 0; }
int lrf_get_boxes(lrf_ctx *h, int32_t *xyz)
{
  NEED_GRID(h);
  if (h->dense) {
    const int n = h->gs - 1;
    size_t i = 0;
    for (int z = 0; z < n; z++)
      for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++, i++) { xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z; }   // id = x + y(gs-1) + z(gs-1)^2, example_tracer.h:77
    return 0;
  }
  if (h->n_boxes) CK(cudaMemcpyAsync(xyz, h->boxes, h->n_boxes * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int lrf_set_view_inv(lrf_ctx *h, const float m[16]) { NEED(h); std::memcpy(h->view_inv, m, 64); return 0; }
int lrf_set_proj_inv(lrf_ctx *h, const float m[16]) { NEED(h); std::memcpy(h->proj_inv, m, 64); return 0; }

// ---- forward ----
int lrf_render_dev_async(lrf_ctx *h, uint32_t *out_dev, uint32_t w, uint32_t hgt, int grey)
{
  NEED_GRID(h);
  if (w == 0 || hgt == 0) return 0;
  if (ensure_fb(h, (size_t)w * hgt)) return 1;
  if (render_dispatch(h, out_dev, w, hgt, grey, nullptr)) return 1;
  CK(cudaGetLastError());
  return 0;
}
// n_views frames in one call: frame v (w * hgt pixels) of out_dev is the render with view matrix view_inv_host[16 v .. 16 v + 15] (the
// projection matrix, grid and options are the context's; the context's own view matrix is left as it was).  ONE launch of the forward
// kernel over all views (no idle SMs between views); only the opt-in staged kernel is run view by view.
// The matrices are copied before the call returns; the frames are complete when the stream reaches this point.
int lrf_render_views_dev_async(lrf_ctx *h, uint32_t n_views, const float *view_inv_host, uint32_t *out_dev, uint32_t w, uint32_t hgt, int grey)
{
  NEED_GRID(h);
  if (n_views == 0 || w == 0 || hgt == 0) return 0;
  if (!view_inv_host || !out_dev) return fail("null view matrices / frame buffer");
  const size_t frame = (size_t)w * hgt;
  // one launch: every configuration the gather kernels render (not the opt-in staged kernel, whose tensor maps are per view)
  const bool one_launch = !(h->staged && !grey) && frame * n_views < (1ull << 32) && (uint64_t)n_views * (((w + 7) / 8) * (uint64_t)((hgt + 3) / 4)) < (1ull << 31);
  if (!one_launch) {
    float keep[16];
    std::memcpy(keep, h->view_inv, 64);
    if (ensure_fb(h, frame)) return 1;
    int rc = 0;
    for (uint32_t v = 0; v < n_views && !rc; v++) {
      std::memcpy(h->view_inv, view_inv_host + 16 * (size_t)v, 64);
      rc = render_dispatch(h, out_dev + frame * v, w, hgt, grey, nullptr);
    }
    std::memcpy(h->view_inv, keep, 64);
    if (rc) return 1;
    CK(cudaGetLastError());
    return 0;
  }
  if (h->views_cap < n_views) {
    cudaFree(h->views_dev);
    h->views_dev = nullptr; h->views_cap = 0;
    CK(cudaMalloc(&h->views_dev, (size_t)n_views * 64));
    h->views_cap = n_views;
  }
  if (ensure_fb(h, frame * n_views)) return 1;
  // pageable host memory: the copy is staged before cudaMemcpyAsync returns, so the caller may reuse the matrices right away
  CK(cudaMemcpyAsync(h->views_dev, view_inv_host, (size_t)n_views * 64, cudaMemcpyHostToDevice, h->stream));
  SceneDev S = scene_of(h);
  if (h->baked && h->dense && !grey) {   // as in render_dispatch
    const int r = ensure_records(h);
    if (r == 1) return 1;
    if (r == 0) S.rec = h->rec;
  }
  launch_render_views(S, h->views_dev, n_views, out_dev, w, hgt, grey, h->fb, h->fb + kFbHeader, h->stream, h->lean);
  h->launches += 2;   // render_lean_kernel / render_image_kernel (all views) + render_fallback_kernel
  h->last_render_staged = 0;
  CK(cudaGetLastError());
  return 0;
}
// rows [y0, y1) of the frame into the full-frame buffer out_dev (latency mode: one frame split over several GPUs in row bands, the
// way the reference splits the rows of a frame over its OpenMP threads, ref example_tracer.cpp:630); gather kernel
int lrf_render_rows_dev_async(lrf_ctx *h, uint32_t *out_dev, uint32_t w, uint32_t hgt, uint32_t y0, uint32_t y1, int grey)
{
  NEED_GRID(h);
  if (w == 0 || hgt == 0 || y0 >= y1) return 0;
  if (y1 > hgt) return fail("row band exceeds the frame");
  if (ensure_fb(h, (size_t)w * hgt)) return 1;
  launch_render_image(scene_of(h), out_dev, w, hgt, grey, h->fb, h->fb + kFbHeader, h->stream, y0, y1, h->lean);
  h->launches += 2;
  h->last_render_staged = 0;
  CK(cudaGetLastError());
  return 0;
}
// parity tap of the staged kernel: frame + per-pixel shaded-sample counts (0xffffffff where the exact kernel rendered the pixel), host buffers
int lrf_render_staged_debug(lrf_ctx *h, uint32_t *out_host, uint32_t *n_shaded_host, uint32_t w, uint32_t hgt)
{
  NEED_GRID(h);
  if (w == 0 || hgt == 0) return 0;
  const size_t bytes = (size_t)w * hgt * 4;
  if (ensure_stage(h, 2 * bytes)) return 1;
  if (ensure_fb(h, (size_t)w * hgt)) return 1;
  uint32_t *frame = (uint32_t *)h->stage, *tap = frame + (size_t)w * hgt;
  const int was = h->staged;
  h->staged = 1;
  const int rc = render_dispatch(h, frame, w, hgt, 0, tap);
  h->staged = was;
  if (rc) return 1;
  CK(cudaGetLastError());
  if (out_host) CKD(h, cudaMemcpyAsync(out_host, frame, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (n_shaded_host) CKD(h, cudaMemcpyAsync(n_shaded_host, tap, bytes, cudaMemcpyDeviceToHost, h->stream));
  return sync_with_diag(h);
}
// parity tap of the register-window march (lrf_lean.cuh): frame + per-pixel shaded-sample count (0xffffffff where the pixel went to the
// generic fallback) + final transmittance, host buffers.  Dense colour grids only.
int lrf_render_lean_debug(lrf_ctx *h, uint32_t *out_host, uint32_t *n_shaded_host, float *thr_host, uint32_t w, uint32_t hgt)
{
  NEED_GRID(h);
  if (w == 0 || hgt == 0) return 0;
  if (!h->dense) return fail("lrf_render_lean_debug: dense box sets only");
  const size_t bytes = (size_t)w * hgt * 4;
  if (ensure_stage(h, 3 * bytes)) return 1;
  if (ensure_fb(h, (size_t)w * hgt)) return 1;
  uint32_t *frame = (uint32_t *)h->stage, *tap = frame + (size_t)w * hgt;
  float *thr = (float *)(tap + (size_t)w * hgt);
  launch_render_image(scene_of(h), frame, w, hgt, 0, h->fb, h->fb + kFbHeader, h->stream, 0, 0xffffffffu, 1, tap, thr);
  h->launches += 2;
  CK(cudaGetLastError());
  if (out_host) CK(cudaMemcpyAsync(out_host, frame, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (n_shaded_host) CK(cudaMemcpyAsync(n_shaded_host, tap, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (thr_host) CK(cudaMemcpyAsync(thr_host, thr, bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
// counters of the staged render since the last reset: [0] rays re-rendered by the exact kernel because the throughput break was too close
// to call, [1] other rays sent there, [2] tiles whose footprint did not fit the staged box, [3] 1 if the last render used the staged kernel
int lrf_render_stats(lrf_ctx *h, uint32_t out[4], int reset)
{
  NEED(h);
  out[0] = out[1] = out[2] = 0;
  out[3] = (uint32_t)h->last_render_staged;
  if (!h->st_stats) return 0;
  CK(cudaMemcpyAsync(out, h->st_stats, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  if (reset) CK(cudaMemsetAsync(h->st_stats, 0, 4 * sizeof(uint32_t), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
// for callers that write the parameters through lrf_param_device_ptr / peer memory: derived copies (the render records) are stale now
int lrf_params_touched(lrf_ctx *h) { NEED(h); params_changed(h); return 0; }
// n_views frames into HOST memory (frame v at out_host + v * w * hgt), synchronous: the views are rendered in up to four batches and the
// finished batch travels to the host on the second stream while the next one is rendered (pinned out_host: the copies overlap the
// kernels; pageable: correct, not overlapped).
int lrf_render_views(lrf_ctx *h, uint32_t n_views, const float *view_inv_host, uint32_t *out_host, uint32_t w, uint32_t hgt, int grey)
{
  NEED_GRID(h);
  if (n_views == 0 || w == 0 || hgt == 0) return 0;
  if (!view_inv_host || !out_host) return fail("null view matrices / frame buffer");
  const size_t frame = (size_t)w * hgt;
  if (ensure_stage(h, frame * n_views * 4)) return 1;
  const uint32_t chunk = (n_views + 3) / 4;
  cudaEvent_t done[4] = {nullptr, nullptr, nullptr, nullptr};
  int nchunk = 0, rc = 0;
  cudaEventRecord(h->ev0, h->stream);
  for (uint32_t v0 = 0; v0 < n_views && !rc; v0 += chunk, nchunk++) {
    const uint32_t nv = std::min(chunk, n_views - v0);
    uint32_t *dev = (uint32_t *)h->stage + frame * v0;
    rc = lrf_render_views_dev_async(h, nv, view_inv_host + 16 * (size_t)v0, dev, w, hgt, grey);
    if (rc) break;
    if (cudaEventCreateWithFlags(&done[nchunk], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(done[nchunk], h->stream) != cudaSuccess ||
        cudaStreamWaitEvent(h->side, done[nchunk], 0) != cudaSuccess ||
        cudaMemcpyAsync(out_host + frame * v0, dev, frame * nv * 4, cudaMemcpyDeviceToHost, h->side) != cudaSuccess)
      rc = fail(std::string("lrf_render_views: ") + cudaGetErrorString(cudaGetLastError()));
  }
  cudaEventRecord(h->ev1, h->stream);
  cudaEventRecord(h->ev_join, h->side);
  cudaStreamWaitEvent(h->stream, h->ev_join, 0);
  const int rs = sync_with_diag(h);
  for (cudaEvent_t e : done)
    if (e) cudaEventDestroy(e);
  if (rc || rs) return 1;
  cudaEventElapsedTime(&h->t_march, h->ev0, h->ev1);
  return 0;
}
int lrf_render(lrf_ctx *h, uint32_t *out_host, uint32_t w, uint32_t hgt, int grey)
{
  NEED_GRID(h);
  if (w == 0 || hgt == 0) return 0;
  const size_t bytes = (size_t)w * hgt * 4;
  if (ensure_stage(h, bytes)) return 1;
  cudaEventRecord(h->ev0, h->stream);
  if (lrf_render_dev_async(h, (uint32_t *)h->stage, w, hgt, grey)) return 1;
  cudaEventRecord(h->ev1, h->stream);
  CKD(h, cudaMemcpyAsync(out_host, h->stage, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (sync_with_diag(h)) return 1;
  cudaEventElapsedTime(&h->t_march, h->ev0, h->ev1);
  return 0;
}

static int run_debug(lrf_ctx *h, bool image, uint32_t w, uint32_t hgt, size_t n, const float *o, const float *d, int grey, uint32_t *n_hit,
                     uint32_t *n_shaded, float *rgbtc, int32_t *vox, uint32_t keep, uint8_t *touched_dev)
{
  const size_t rays = image ? (size_t)w * hgt : n;
  if (rays == 0) return 0;
  // staging layout: [o | d | n_hit | n_shaded | rgbtc | vox]
  const size_t b_o = image ? 0 : rays * 12, b_nh = rays * 4, b_rg = rays * 20, b_vx = vox ? rays * keep * 12 : 0;
  const size_t total = 2 * b_o + 2 * b_nh + b_rg + b_vx + 256;
  if (ensure_stage(h, total)) return 1;
  char *base = (char *)h->stage;
  float *o_dev = (float *)base, *d_dev = (float *)(base + b_o);
  DebugTaps taps{};
  taps.n_hit = (uint32_t *)(base + 2 * b_o);
  taps.n_shaded = (uint32_t *)(base + 2 * b_o + b_nh);
  taps.rgbtc = (float *)(base + 2 * b_o + 2 * b_nh);
  taps.vox_xyz = vox ? (int32_t *)(base + 2 * b_o + 2 * b_nh + b_rg) : nullptr;
  taps.keep = vox ? keep : 0;
  taps.touched = touched_dev;
  if (!image) {
    CK(cudaMemcpyAsync(o_dev, o, b_o, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_dev, d, b_o, cudaMemcpyHostToDevice, h->stream));
    launch_debug_rays(scene_of(h), rays, o_dev, d_dev, grey, taps, h->force_generic, h->stream);
  } else launch_debug_image(scene_of(h), w, hgt, grey, taps, h->force_generic, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  if (n_hit) CK(cudaMemcpyAsync(n_hit, taps.n_hit, b_nh, cudaMemcpyDeviceToHost, h->stream));
  if (n_shaded) CK(cudaMemcpyAsync(n_shaded, taps.n_shaded, b_nh, cudaMemcpyDeviceToHost, h->stream));
  if (rgbtc) CK(cudaMemcpyAsync(rgbtc, taps.rgbtc, b_rg, cudaMemcpyDeviceToHost, h->stream));
  if (vox) CK(cudaMemcpyAsync(vox, taps.vox_xyz, b_vx, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int lrf_render_debug(lrf_ctx *h, uint32_t w, uint32_t hgt, int grey, uint32_t *n_hit, uint32_t *n_shaded, float *rgbtc, int32_t *vox, uint32_t keep)
{
  NEED_GRID(h);
  return run_debug(h, true, w, hgt, 0, nullptr, nullptr, grey, n_hit, n_shaded, rgbtc, vox, keep, nullptr);
}
int lrf_trace_rays(lrf_ctx *h, size_t n, const float *o, const float *d, int grey, uint32_t *n_hit, uint32_t *n_shaded, float *rgbtc, int32_t *vox,
                   uint32_t keep)
{
  NEED_GRID(h);
  return run_debug(h, false, 0, 0, n, o, d, grey, n_hit, n_shaded, rgbtc, vox, keep, nullptr);
}
int lrf_count_touched_nodes(lrf_ctx *h, uint32_t w, uint32_t hgt, int grey, uint64_t *out)
{
  NEED_GRID(h);
  uint8_t *marks = nullptr;
  unsigned long long *cnt = nullptr;
  CK(cudaMalloc(&marks, h->N));
  CK(cudaMalloc(&cnt, 8));
  CK(cudaMemsetAsync(marks, 0, h->N, h->stream));
  CK(cudaMemsetAsync(cnt, 0, 8, h->stream));
  int rc = run_debug(h, true, w, hgt, 0, nullptr, nullptr, grey, nullptr, nullptr, nullptr, nullptr, 0, marks);
  if (!rc) {
    launch_count_nonzero(marks, h->N, cnt, h->stream);
    h->launches++;
    unsigned long long v = 0;
    cudaMemcpyAsync(&v, cnt, 8, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    *out = v;
  }
  cudaFree(marks); cudaFree(cnt);
  return rc;
}

// ---- backward ----
int lrf_render_grad_dev_async(lrf_ctx *h, uint32_t w, uint32_t hgt, const float *ref_dev, int grey)
{
  NEED_GRID(h);
  if (w == 0 || hgt == 0) return 0;
  if (h->generic_image_grad) {
    launch_grad_image(scene_of(h), w, hgt, ref_dev, grey, nullptr, nullptr, h->stream);
    h->launches++;
  } else {
    if (ensure_fb(h, (size_t)w * hgt)) return 1;
    launch_grad_image(scene_of(h), w, hgt, ref_dev, grey, h->fb, h->fb + kFbHeader, h->stream);
    h->launches += 2;   // grad_image_fast_kernel + the generic kernel over its fallback queue
  }
  CK(cudaGetLastError());
  return 0;
}
int lrf_render_grad(lrf_ctx *h, uint32_t w, uint32_t hgt, const float *ref_host, int grey)
{
  NEED_GRID(h);
  if (w == 0 || hgt == 0) return 0;
  const size_t bytes = (size_t)w * hgt * 16;
  if (ensure_stage(h, bytes)) return 1;
  CK(cudaMemcpyAsync(h->stage, ref_host, bytes, cudaMemcpyHostToDevice, h->stream));
  cudaEventRecord(h->ev0, h->stream);
  if (lrf_render_grad_dev_async(h, w, hgt, (const float *)h->stage, grey)) return 1;
  cudaEventRecord(h->ev1, h->stream);
  CK(cudaStreamSynchronize(h->stream));   // ref buffer is caller-owned and freed right after L1Loss (main.cpp:118)
  cudaEventElapsedTime(&h->t_grad, h->ev0, h->ev1);
  return 0;
}
int lrf_rays_grad_dev_async(lrf_ctx *h, size_t n, const float *o, const float *d, const float *ref, int grey)
{
  NEED_GRID(h);
  if (n == 0) return 0;
  if (ensure_fb(h, n)) return 1;
  // scratch for the per-ray hit lists (pass A -> pass B of the warp kernel); past 512 MB the kernel walks twice instead
  const size_t hb = grad_rays_hit_bytes(n, h->gs, nullptr);
  void *hits = nullptr;
  if (hb <= ((size_t)512 << 20)) {
    if (h->hit_bytes < hb) {
      cudaFree(h->hit_scratch);
      h->hit_scratch = nullptr; h->hit_bytes = 0;
      CK(cudaMalloc(&h->hit_scratch, hb));
      h->hit_bytes = hb;
    }
    hits = h->hit_scratch;
  }
  launch_grad_rays(scene_of(h), n, o, d, ref, grey, h->fb, h->fb + kFbHeader, hits, h->stream, h->wray_parallel);
  h->launches += 2;   // grad_rays_warp_kernel + the generic fallback over its queue
  CK(cudaGetLastError());
  return 0;
}
int lrf_rays_grad(lrf_ctx *h, size_t n, const float *o, const float *d, const float *ref, int grey)
{
  NEED_GRID(h);
  if (n == 0) return 0;
  if (ensure_stage(h, n * 40 + 64)) return 1;
  float *o_dev = (float *)h->stage, *d_dev = o_dev + 3 * n, *r_dev = d_dev + 3 * n;
  // keep the float4 ref 16-byte aligned
  r_dev = (float *)(((uintptr_t)r_dev + 15) & ~(uintptr_t)15);
  CK(cudaMemcpyAsync(o_dev, o, n * 12, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_dev, d, n * 12, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(r_dev, ref, n * 16, cudaMemcpyHostToDevice, h->stream));
  if (lrf_rays_grad_dev_async(h, n, o_dev, d_dev, r_dev, grey)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int lrf_gen_rays_dev_async(lrf_ctx *h, size_t n, const float *view_inv_dev, const uint32_t *picks_dev, uint32_t w, uint32_t hgt, float *o_dev,
                           float *d_dev)
{
  NEED(h);
  launch_gen_rays(h->proj_inv, n, view_inv_dev, picks_dev, w, hgt, o_dev, d_dev, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

// ---- callers either side of the path ----
// model.dat: the raw gs^3 x Cell dump the reference keeps commented in main.cpp:54-56 (load) and :139-141 (save)
int lrf_model_save(lrf_ctx *h, const char *path)
{
  NEED_GRID(h);
  if (!path) return fail("null path");
  std::vector<float> cells(h->N * kCell);
  if (lrf_grid_download(h, cells.data())) return 1;
  FILE *f = std::fopen(path, "wb");
  if (!f) return fail(std::string("cannot open for writing: ") + path);
  const size_t wr = std::fwrite(cells.data(), sizeof(float), cells.size(), f);
  const int rc = std::fclose(f);
  if (wr != cells.size() || rc != 0) return fail(std::string("short write: ") + path);
  return 0;
}
int lrf_model_load(lrf_ctx *h, const char *path)
{
  NEED_GRID(h);
  if (!path) return fail("null path");
  FILE *f = std::fopen(path, "rb");
  if (!f) return fail(std::string("cannot open: ") + path);
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  const size_t want = h->N * kCell * sizeof(float);
  if (sz < 0 || (size_t)sz != want) {
    std::fclose(f);
    return fail(std::string(path) + ": " + std::to_string(sz) + " bytes, but a " + std::to_string(h->gs) + "^3 grid of 52-byte cells is " +
                std::to_string(want) + " (the reference reads grid.size()*sizeof(Cell) bytes, main.cpp:55)");
  }
  std::vector<float> cells(h->N * kCell);
  const size_t rd = std::fread(cells.data(), sizeof(float), cells.size(), f);
  std::fclose(f);
  if (rd != cells.size()) return fail(std::string("short read: ") + path);
  return lrf_grid_upload(h, cells.data());
}

int lrf_targets_resize(lrf_ctx *h, uint32_t n_views, uint32_t w, uint32_t hgt)
{
  NEED(h);
  cudaFree(h->targets);
  h->targets = nullptr;
  h->n_targets = h->tw = h->th = 0;
  const size_t bytes = (size_t)n_views * w * hgt * 16;
  if (bytes) CK(cudaMalloc(&h->targets, bytes));
  h->n_targets = n_views; h->tw = w; h->th = hgt;
  if (!h->gamma_lut) {
    float lut[256];
    const float gamma = 2.2f, scale = 1.0f;                       // stb_image.h:1572 defaults; main.cpp never changes them
    for (int i = 0; i < 256; i++) lut[i] = (float)(powf((float)i / 255.0f, gamma) * scale);   // stb_image.h:1868 as C++ sees it: pow(float,float)
    CK(cudaMalloc(&h->gamma_lut, sizeof(lut)));
    CK(cudaMemcpy(h->gamma_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
  }
  return 0;
}
#define NEED_TARGET(h, view) do { NEED(h); if (!(h)->targets || (view) >= (h)->n_targets) return fail("target slot out of range (call lrf_targets_resize)"); } while (0)
int lrf_target_ingest_rgba8(lrf_ctx *h, uint32_t view, const uint8_t *rgba8_host, int flip_vertical)
{
  NEED_TARGET(h, view);
  const size_t px = (size_t)h->tw * h->th;
  if (ensure_stage(h, px * 4)) return 1;
  CK(cudaMemcpyAsync(h->stage, rgba8_host, px * 4, cudaMemcpyHostToDevice, h->stream));
  launch_ingest_rgba8((const uint8_t *)h->stage, h->targets + (size_t)view * px * 4, h->tw, h->th, flip_vertical, h->gamma_lut, h->stream);
  h->launches++;
  CK(cudaStreamSynchronize(h->stream));   // the caller's buffer and the staging area are free again
  return 0;
}
int lrf_target_upload_rgbaf(lrf_ctx *h, uint32_t view, const float *rgba_host)
{
  NEED_TARGET(h, view);
  const size_t px = (size_t)h->tw * h->th;
  CK(cudaMemcpyAsync(h->targets + (size_t)view * px * 4, rgba_host, px * 16, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int lrf_target_download(lrf_ctx *h, uint32_t view, float *rgba_host)
{
  NEED_TARGET(h, view);
  const size_t px = (size_t)h->tw * h->th;
  CK(cudaMemcpyAsync(rgba_host, h->targets + (size_t)view * px * 4, px * 16, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
float *lrf_target_device_ptr(lrf_ctx *h, uint32_t view)
{
  if (!h || !h->targets || view >= h->n_targets) return nullptr;
  return h->targets + (size_t)view * h->tw * h->th * 4;
}
int lrf_render_grad_target_async(lrf_ctx *h, uint32_t view, int grey)
{
  NEED_TARGET(h, view);
  return lrf_render_grad_dev_async(h, h->tw, h->th, lrf_target_device_ptr(h, view), grey);
}
int lrf_gather_targets_dev_async(lrf_ctx *h, size_t n, const uint32_t *picks_dev, float *ref_rgba_dev)
{
  NEED(h);
  if (!h->targets) return fail("no resident targets (call lrf_targets_resize)");
  launch_gather_targets(n, picks_dev, h->targets, (size_t)h->tw * h->th, ref_rgba_dev, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

// BMP: LiteImage::SaveBMP / WriteBMP, ref external/LiteMath/Image2d.cpp:488-533 (header bytes and record order restated)
static int write_bmp_file(const char *path, const uint8_t *records, int width, int height)
{
  const int paddedsize = width * height * 3;
  unsigned char fh[14] = {'B', 'M', 0, 0, 0, 0, 0, 0, 0, 0, 54, 0, 0, 0};
  unsigned char ih[40] = {40, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 24, 0};
  for (int b = 0; b < 4; b++) {
    fh[2 + b] = (unsigned char)(paddedsize >> (8 * b));    // the reference stores the pixel-array size here, not the file size
    ih[4 + b] = (unsigned char)(width >> (8 * b));
    ih[8 + b] = (unsigned char)(height >> (8 * b));
  }
  FILE *f = std::fopen(path, "wb");
  if (!f) return fail(std::string("cannot open for writing: ") + path);
  size_t wr = std::fwrite(fh, 1, 14, f) + std::fwrite(ih, 1, 40, f) + std::fwrite(records, 1, (size_t)paddedsize, f);
  const int rc = std::fclose(f);
  if (wr != (size_t)paddedsize + 54 || rc != 0) return fail(std::string("short write: ") + path);
  return 0;
}
int lrf_save_bmp(const char *path, const uint32_t *pixels_host, int width, int height)
{
  if (!path || !pixels_host || width < 0 || height < 0) return fail("bad argument");
  std::vector<uint8_t> rec((size_t)width * height * 3);
  for (size_t i = 0; i < (size_t)width * height; i++) {
    rec[3 * i] = (uint8_t)((pixels_host[i] & 0x00FF0000u) >> 16);
    rec[3 * i + 1] = (uint8_t)((pixels_host[i] & 0x0000FF00u) >> 8);
    rec[3 * i + 2] = (uint8_t)(pixels_host[i] & 0x000000FFu);
  }
  return write_bmp_file(path, rec.data(), width, height);
}
int lrf_render_bmp(lrf_ctx *h, const char *path, uint32_t w, uint32_t hgt, int grey)
{
  NEED_GRID(h);
  if (!path) return fail("null path");
  const size_t px = (size_t)w * hgt, frame = (px * 4 + 255) & ~(size_t)255;
  if (ensure_stage(h, frame + px * 3 + 16)) return 1;
  cudaEventRecord(h->ev0, h->stream);
  if (px && lrf_render_dev_async(h, (uint32_t *)h->stage, w, hgt, grey)) return 1;
  cudaEventRecord(h->ev1, h->stream);
  launch_pack_bmp((const uint32_t *)h->stage, (uint8_t *)h->stage + frame, px, h->stream);
  h->launches++;
  std::vector<uint8_t> rec(px * 3);
  if (px) CK(cudaMemcpyAsync(rec.data(), (uint8_t *)h->stage + frame, px * 3, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->t_march, h->ev0, h->ev1);
  return write_bmp_file(path, rec.data(), (int)w, (int)hgt);
}

// ---- TV / Adam ----
static int ensure_tv_tmp(lrf_ctx *h)
{
  if (!h->tv_tmp) CK(cudaMalloc(&h->tv_tmp, 2 * h->N * sizeof(float)));
  return 0;
}
int lrf_tv_grad(lrf_ctx *h)
{
  NEED_GRID(h);
  if (ensure_tv_tmp(h)) return 1;
  cudaEventRecord(h->ev0, h->stream);
  tv_dispatch(h, 0, h->gs, 0, 0, true, h->stream);
  h->launches += (h->gs >= 8) ? 1 : 3;   // one TV kernel (z-sweep, or 8^3 tiles below 64^3), stage-by-stage below 8^3
  cudaEventRecord(h->ev1, h->stream);
  CK(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->t_tv, h->ev0, h->ev1);
  return 0;
}
int lrf_adam_init(lrf_ctx *h)
{
  NEED_GRID(h);
  h->lr = 0.15f; h->beta1 = 0.9f; h->beta2 = 0.999f; h->eps = 1e-8;
  const size_t bytes = h->N * kCell * sizeof(float);
  CK(cudaMemsetAsync(h->V, 0, bytes, h->stream));
  CK(cudaMemsetAsync(h->S, 0, bytes, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int lrf_set_adam_hparams(lrf_ctx *h, float lr, float b1, float b2, float eps)
{
  NEED(h);
  h->lr = lr; h->beta1 = b1; h->beta2 = b2; h->eps = eps;
  return 0;
}
static void adam_launch(lrf_ctx *h, int iter, int which, int zero, cudaStream_t st = nullptr)
{
  const double b1 = std::pow(h->beta1, iter + 1), b2 = std::pow(h->beta2, iter + 1);   // example_tracer.h:152-153 (double)
  size_t off = 0, count = h->N * kCell;
  if (which == 1) count = h->N;                            // density plane
  else if (which == 2) { off = h->N; count = h->N * 12; }  // everything but density (i % CellSize != 0)
  launch_adam(h->params, h->grad, h->V, h->S, off, count, h->lr, h->beta1, h->beta2, h->eps, 1 - b1, 1 - b2, zero, st ? st : h->stream);
  params_changed(h);
  h->launches++;
}
int lrf_adam_step(lrf_ctx *h, int iter, int which)
{
  NEED_GRID(h);
  if (which < 0 || which > 2) return fail("which must be 0 (all), 1 (density) or 2 (SH/RGB)");
  cudaEventRecord(h->ev0, h->stream);
  adam_launch(h, iter, which, 0);
  cudaEventRecord(h->ev1, h->stream);
  CK(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->t_adam, h->ev0, h->ev1);
  return 0;
}
int lrf_train_percell_async(lrf_ctx *h, int iter)
{
  NEED_GRID(h);
  if (ensure_tv_tmp(h)) return 1;
  // Measured at 256^3 (tools/time_percell.py): the 8^3-tile TV kernel beside Adam bought nothing (1.41 ms either way: it needs the whole
  // GPU to be fast); the z-sweep kernel does — 1.35 ms in sequence, 1.27 ms overlapped, 1.24 ms with its grid capped at 2 blocks per SM.
  if (h->percell_overlap && h->gs >= 64) {
    // {TV -> Adam(density plane)} beside Adam(colour planes) on the second stream: the z-sweep TV kernel is light enough on the SMs
    // (32 registers, 18 KB of shared memory, no index arithmetic) to hide behind the HBM-bound colour pass
    CK(cudaEventRecord(h->ev_fork, h->stream));
    CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    if (h->percell_overlap == 2) adam_launch(h, iter, 2, 1);          // order experiment: the colour pass queued first
    tv_dispatch(h, 0, h->gs, 0, 0, true, h->side, h->tv_blocks_per_sm);
    h->launches++;
    adam_launch(h, iter, 1, 1, h->side);
    CK(cudaEventRecord(h->ev_join, h->side));
    if (h->percell_overlap != 2) adam_launch(h, iter, 2, 1);
    CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    CK(cudaGetLastError());
    return 0;
  }
  tv_dispatch(h, 0, h->gs, 0, 0, true, h->stream);
  h->launches += (h->gs >= 8) ? 1 : 3;
  adam_launch(h, iter, 0, 1);
  CK(cudaGetLastError());
  return 0;
}

// training iteration, first half: the ray backward of a packed batch and — concurrently, on a second stream — the TV gradient
// (it depends on the parameters only; both scatter into the gradient planes with red.add).  With a peer group attached the TV
// covers the own z-slab only.  Joined on the context stream before returning control to it.
int lrf_rays_grad_tv_dev_async(lrf_ctx *h, size_t n, const float *o, const float *d, const float *ref, int grey)
{
  NEED_GRID(h);
  if (h->gs < 8) {   // tiny grids use the staged TV kernels (not atomic): run one after the other
    if (lrf_rays_grad_dev_async(h, n, o, d, ref, grey)) return 1;
    if (ensure_tv_tmp(h)) return 1;
    launch_tv_grad(h->gs, h->params, h->grad, h->tv_tmp, h->tv_tmp + h->N, h->stream);
    h->launches += 3;
    return 0;
  }
  CK(cudaEventRecord(h->ev_fork, h->stream));
  CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  int z0 = 0, z1 = h->gs;
  if (h->peer_world > 1) { const int zs = h->gs / h->peer_world; z0 = zs * h->peer_rank; z1 = z0 + zs; }
  tv_dispatch(h, z0, z1, 1, 0, false, h->side);
  h->launches++;
  CK(cudaEventRecord(h->ev_join, h->side));
  if (lrf_rays_grad_dev_async(h, n, o, d, ref, grey)) return 1;
  CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  CK(cudaGetLastError());
  return 0;
}
// second half on one GPU: optimizerStep(iter) + gradient zeroing for the next iteration
int lrf_adam_zero_async(lrf_ctx *h, int iter)
{
  NEED_GRID(h);
  adam_launch(h, iter, 0, 1);
  CK(cudaGetLastError());
  return 0;
}

// ---- multi-GPU exchange over peer memory ----
int lrf_ipc_export(lrf_ctx *h, int which, void *handle64)
{
  NEED_GRID(h);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t hd;
  CK(cudaIpcGetMemHandle(&hd, which == 2 ? (void *)h->flags : which ? (void *)h->grad : (void *)h->params));
  std::memcpy(handle64, &hd, 64);
  return 0;
}
int lrf_peer_attach_ipc(lrf_ctx *h, int world, int rank, const void *param_handles, const void *grad_handles)
{
  NEED_GRID(h);
  if (peer_check(h, world, rank)) return 1;
  peer_detach(h);
  for (int r = 0; r < world; r++) {
    if (r == rank) { h->peer_param[r] = h->params; h->peer_grad[r] = h->grad; continue; }
    cudaIpcMemHandle_t hp, hg;
    std::memcpy(&hp, (const char *)param_handles + 64 * r, 64);
    std::memcpy(&hg, (const char *)grad_handles + 64 * r, 64);
    CK(cudaIpcOpenMemHandle((void **)&h->peer_param[r], hp, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle((void **)&h->peer_grad[r], hg, cudaIpcMemLazyEnablePeerAccess));
  }
  h->peer_world = world; h->peer_rank = rank; h->peer_ipc = true;
  return 0;
}
int lrf_peer_attach_local(lrf_ctx *h, int world, int rank, lrf_ctx *const *ctxs)
{
  NEED_GRID(h);
  if (peer_check(h, world, rank)) return 1;
  peer_detach(h);
  for (int r = 0; r < world; r++) {
    lrf_ctx *o = ctxs[r];
    if (!o || !o->params || o->gs != h->gs) return fail("peer group: every context needs a grid of the same size");
    if (o->device != h->device) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, h->device, o->device));
      if (!can) return fail("peer group: device cannot access its peer");
      cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
    h->peer_param[r] = o->params; h->peer_grad[r] = o->grad;
    h->peer_flags[r] = o->flags;
  }
  if (ctxs[rank] != h) return fail("peer group: ctxs[rank] must be this context");
  h->peer_world = world; h->peer_rank = rank; h->peer_ipc = false;
  return 0;
}
int lrf_peer_detach(lrf_ctx *h) { NEED(h); peer_detach(h); return 0; }
// barrier flags of the group (after lrf_peer_attach_*): flag_handles = world x 64 bytes from lrf_ipc_export(which = 2), or
// NULL for a group attached with lrf_peer_attach_local (ctxs give the arrays directly)
int lrf_peer_attach_flags_ipc(lrf_ctx *h, const void *flag_handles)
{
  NEED(h);
  if (h->peer_world < 1 || !h->peer_ipc) return fail("attach the group with lrf_peer_attach_ipc or lrf_mc_prepare first");
  for (int r = 0; r < h->peer_world; r++) {
    if (r == h->peer_rank) { h->peer_flags[r] = h->flags; continue; }
    cudaIpcMemHandle_t hf;
    std::memcpy(&hf, (const char *)flag_handles + 64 * r, 64);
    CK(cudaIpcOpenMemHandle((void **)&h->peer_flags[r], hf, cudaIpcMemLazyEnablePeerAccess));
  }
  return 0;
}
int lrf_peer_barrier_async(lrf_ctx *h)
{
  NEED(h);
  if (h->peer_world < 1 || !h->peer_flags[h->peer_rank]) return fail("no barrier flags attached (lrf_peer_attach_flags_ipc / lrf_peer_attach_local)");
  h->peer_epoch++;
  launch_peer_barrier(h->peer_world, h->peer_rank, h->peer_flags, h->peer_epoch, h->flags + 16, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}
// 1 if a barrier of this context gave up waiting for a peer (2 s) since the last call; synchronises the stream
int lrf_peer_barrier_timed_out(lrf_ctx *h)
{
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  uint32_t v = 0;
  cudaMemcpy(&v, h->flags + 16, 4, cudaMemcpyDeviceToHost);
  if (v) cudaMemset(h->flags + 16, 0, 4);
  return v ? 1 : 0;
}
#define NEED_PEERS(h) do { NEED_GRID(h); if ((h)->peer_world < 1) return fail("no peer group attached (lrf_peer_attach_ipc / _local)"); } while (0)
int lrf_tv_grad_slab_async(lrf_ctx *h)
{
  NEED_PEERS(h);
  const int zs = h->gs / h->peer_world;
  tv_dispatch(h, zs * h->peer_rank, zs * (h->peer_rank + 1), 0, 0, false, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}
int lrf_adam_peer_async(lrf_ctx *h, int iter)
{
  NEED_PEERS(h);
  const double b1 = std::pow(h->beta1, iter + 1), b2 = std::pow(h->beta2, iter + 1);   // example_tracer.h:152-153 (double)
  const size_t n = (size_t)h->peer_world, r = (size_t)h->peer_rank;
  // own slab of the density plane and of the colour plane (contiguous z-slabs of both)
  params_changed(h);   // every rank runs this pass and every replica is rewritten by it
  if (h->mc) {
    if (h->mc_state != 4) return fail("multicast group not bound yet (lrf_mc_bind)");
    float *mcp = reinterpret_cast<float *>(h->mc_va), *mcg = reinterpret_cast<float *>(h->mc_va + h->mc_bytes);
    launch_adam_mc(mcg, mcp, h->params, h->V, h->S, h->N / n * r, h->N / n, h->lr, h->beta1, h->beta2, h->eps, 1 - b1, 1 - b2, h->stream, h->mc_unroll, h->mc_blocks_per_sm);
    launch_adam_mc(mcg, mcp, h->params, h->V, h->S, h->N + h->N * 12 / n * r, h->N * 12 / n, h->lr, h->beta1, h->beta2, h->eps, 1 - b1, 1 - b2, h->stream, h->mc_unroll, h->mc_blocks_per_sm);
    h->launches += 2;
    CK(cudaGetLastError());
    return 0;
  }
  launch_adam_peer(h->peer_world, h->peer_rank, h->peer_grad, h->peer_param, h->V, h->S, h->N / n * r, h->N / n, h->lr, h->beta1, h->beta2, h->eps,
                   1 - b1, 1 - b2, h->stream);
  launch_adam_peer(h->peer_world, h->peer_rank, h->peer_grad, h->peer_param, h->V, h->S, h->N + h->N * 12 / n * r, h->N * 12 / n, h->lr, h->beta1,
                   h->beta2, h->eps, 1 - b1, 1 - b2, h->stream);
  h->launches += 2;
  CK(cudaGetLastError());
  return 0;
}
int lrf_train_percell_peer_async(lrf_ctx *h, int iter)
{
  if (lrf_tv_grad_slab_async(h)) return 1;
  return lrf_adam_peer_async(h, iter);
}
int lrf_zero_grad_async(lrf_ctx *h)
{
  NEED_GRID(h);
  launch_fill(h->grad, h->N * kCell, 0.0f, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

// ---- level transition ----
int lrf_upsample_grid(lrf_ctx *h)
{
  NEED_GRID(h);
  if (h->gs * 2 > 1024) return fail("upsampled grid would exceed 1024^3");
  const int old_gs = h->gs;
  float *old_params = h->params;
  const bool was_dense = h->dense;
  int32_t *old_boxes = h->boxes;
  const size_t old_nb = h->n_boxes;
  // detach what must survive alloc_grid
  h->params = nullptr; h->boxes = nullptr;
  if (alloc_grid(h, old_gs * 2)) { cudaFree(old_params); cudaFree(old_boxes); return 1; }
  launch_upsample_grid(old_gs, old_params, h->params, h->stream);
  params_changed(h);
  h->launches++;
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(old_params);
  // The reference's UpsampleGrid leaves `boxes` untouched (old-level boxes, twice the voxel size) until
  // UpsampleOctree runs; rendering in between is not meaningful there either.  Keep the old list for
  // lrf_upsample_boxes and mark the set as pending.
  if (was_dense) {
    // materialise the dense old-level list so that lrf_upsample_boxes can subdivide it
    const int n = old_gs - 1;
    const size_t nb = (size_t)n * n * n;
    CK(cudaMalloc(&h->boxes, (nb ? nb : 1) * 3 * sizeof(int32_t)));
    launch_dense_boxes(n, h->boxes, h->stream);
    h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    h->n_boxes = nb;
  } else { h->boxes = old_boxes; h->n_boxes = old_nb; }
  h->dense = false;  // old-level boxes: maps are rebuilt by lrf_upsample_boxes
  return 0;
}

static int compact_boxes(lrf_ctx *h, bool children)
{
  const size_t n_old = h->n_boxes, n_flags = children ? n_old * 8 : n_old;
  if (n_flags == 0) { int32_t *e = nullptr; CK(cudaMalloc(&e, 12)); return install_boxes_dev(h, e, 0); }
  uint32_t *flags = nullptr, *offs = nullptr, *tmp = nullptr;
  CK(cudaMalloc(&flags, n_flags * 4));
  CK(cudaMalloc(&offs, n_flags * 4));
  CK(cudaMalloc(&tmp, scan_tmp_elems(n_flags) * 4));
  const SceneDev S = scene_of(h);
  if (children) launch_box_child_flags(S, h->boxes, n_old, flags, h->stream);
  else launch_box_keep_flags(S, h->boxes, n_old, flags, h->stream);
  launch_exclusive_scan(flags, offs, n_flags, tmp, h->stream);
  h->launches += 4;
  uint32_t last_off = 0, last_flag = 0;
  CK(cudaMemcpyAsync(&last_off, offs + n_flags - 1, 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&last_flag, flags + n_flags - 1, 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const size_t n_new = (size_t)last_off + last_flag;
  int32_t *nb = nullptr;
  CK(cudaMalloc(&nb, (n_new ? n_new : 1) * 3 * sizeof(int32_t)));
  if (children) launch_box_compact_children(h->boxes, n_old, flags, offs, nb, h->stream);
  else launch_box_compact(h->boxes, n_old, flags, offs, nb, h->stream);
  h->launches++;
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(flags); cudaFree(offs); cudaFree(tmp);
  return install_boxes_dev(h, nb, n_new);
}
int lrf_upsample_boxes(lrf_ctx *h)
{
  NEED_GRID(h);
  if (h->dense) return fail("lrf_upsample_boxes must follow lrf_upsample_grid (main.cpp:128-131)");
  return compact_boxes(h, true);
}
int lrf_rebuild_boxes(lrf_ctx *h)
{
  NEED_GRID(h);
  if (h->dense) {   // materialise the InitGrid list on the device (133 M boxes at 512^3: no host round trip)
    const size_t n = h->n_boxes;
    CK(cudaMalloc(&h->boxes, (n ? n : 1) * 3 * sizeof(int32_t)));
    launch_dense_boxes(h->gs - 1, h->boxes, h->stream);
    h->launches++;
    h->dense = false;   // compact_boxes reads h->boxes and installs the filtered list (+ occupancy maps)
  }
  return compact_boxes(h, false);
}

// ---- plumbing ----
int lrf_last_time_ms(lrf_ctx *h, const char *name, float out[4])
{
  NEED(h);
  const std::string s(name ? name : "");
  if (s == "RayMarch") out[0] = h->t_march;
  else if (s == "RayMarchGrad") out[0] = h->t_grad;
  else if (s == "TVGrad") out[0] = h->t_tv;
  else if (s == "Adam") out[0] = h->t_adam;
  else return fail("unknown timer name");
  return 0;
}
int lrf_timer_start(lrf_ctx *h) { NEED(h); CK(cudaEventRecord(h->tm0, h->stream)); return 0; }
int lrf_timer_stop_ms(lrf_ctx *h, float *ms)
{
  NEED(h);
  CK(cudaEventRecord(h->tm1, h->stream));
  CKD(h, cudaEventSynchronize(h->tm1));
  CK(cudaEventElapsedTime(ms, h->tm0, h->tm1));
  return 0;
}
float *lrf_param_device_ptr(lrf_ctx *h) { return h ? h->params : nullptr; }   // writers must call lrf_params_touched afterwards
float *lrf_grad_device_ptr(lrf_ctx *h) { return h ? h->grad : nullptr; }
size_t lrf_param_count(lrf_ctx *h) { return h ? h->N * kCell : 0; }
void *lrf_stream(lrf_ctx *h) { return h ? (void *)h->stream : nullptr; }
int lrf_use_stream(lrf_ctx *h, void *s)
{
  NEED(h);
  CK(cudaStreamSynchronize(h->stream));
  if (h->own_stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)s;
  h->own_stream = false;
  return 0;
}
int lrf_device_alloc(lrf_ctx *h, size_t bytes, void **out) { NEED(h); CK(cudaMalloc(out, bytes)); return 0; }
int lrf_device_free(lrf_ctx *h, void *p) { NEED(h); CK(cudaFree(p)); return 0; }
int lrf_host_alloc_pinned(size_t bytes, void **out) { CK(cudaMallocHost(out, bytes)); return 0; }
int lrf_host_free_pinned(void *p) { CK(cudaFreeHost(p)); return 0; }
int lrf_memcpy_h2d_async(lrf_ctx *h, void *dst, const void *src, size_t bytes) { NEED(h); CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream)); return 0; }
int lrf_memcpy_d2h_async(lrf_ctx *h, void *dst, const void *src, size_t bytes) { NEED(h); CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream)); return 0; }
uint64_t lrf_launch_count(lrf_ctx *h) { return h ? h->launches : 0; }
int lrf_set_option(lrf_ctx *h, const char *name, int value)
{
  NEED(h);
  const std::string s(name ? name : "");
  if (s == "force_generic_march") h->force_generic = value ? 1 : 0;
  else if (s == "generic_image_grad") h->generic_image_grad = value ? 1 : 0;
  else if (s == "staged_render") h->staged = value ? 1 : 0;
  else if (s == "lean_render") h->lean = value ? 1 : 0;
  else if (s == "wray_parallel") h->wray_parallel = value ? 1 : 0;
  else if (s == "mc_unroll") h->mc_unroll = value;
  else if (s == "mc_blocks_per_sm") h->mc_blocks_per_sm = value;
  else if (s == "l2_persist_density") {
    // experiment (north_star: "L2-persisting windows"): value = percent of the density plane (0 = off) marked persisting in L2 for the
    // kernels of this context's stream; the plane is 4 N bytes (67 MB at 256^3, the L2 is 126 MB).  Measured: DESIGN.md 4-K1c.
    NEED_GRID(h);
    cudaStreamAttrValue av{};
    if (value > 0) {
      int max_persist = 0, max_win = 0;
      CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device));
      CK(cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, h->device));
      size_t bytes = h->N * sizeof(float);
      if (bytes > (size_t)max_win) bytes = (size_t)max_win;
      CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist));
      av.accessPolicyWindow.base_ptr = h->params;
      av.accessPolicyWindow.num_bytes = bytes;
      av.accessPolicyWindow.hitRatio = std::min(1.0f, (float)value / 100.0f * (float)max_persist / (float)bytes);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      if (std::getenv("LRF_DEBUG")) std::fprintf(stderr, "[lrf] L2 window: %zu bytes, hit ratio %.2f (persisting L2 max %d MB, window max %d MB)\n", bytes, av.accessPolicyWindow.hitRatio, max_persist >> 20, max_win >> 20);
    } else {
      av.accessPolicyWindow.num_bytes = 0;
      av.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
      av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    CK(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av));
    if (value <= 0) CK(cudaCtxResetPersistingL2Cache());
  }
  else if (s == "baked_records") h->baked = value ? 1 : 0;
  else if (s == "tv_sweep") h->tv_sweep = value ? 1 : 0;
  else if (s == "percell_overlap") h->percell_overlap = value;
  else if (s == "tv_blocks_per_sm") h->tv_blocks_per_sm = value;
  else if (s == "staged_warps") { if (value != 4 && value != 8) return fail("staged_warps must be 4 or 8"); h->staged_nw = value; }
  else if (s == "staged_box_w") h->staged_box_w = value;
  else if (s == "staged_box_h") h->staged_box_h = value;
  else return fail("unknown option: " + s);
  return 0;
}

}  // extern "C"
