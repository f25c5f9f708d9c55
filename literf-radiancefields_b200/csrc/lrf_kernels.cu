// lrf_kernels.cu — hand-written sm_100a kernels of the LiteRF ray-march hot path.
//
//   K1 render / debug   : ref kernel2D_RayMarch           example_tracer.cpp:627-663
//   K2 grad             : ref kernel2D_RayMarchGrad       example_tracer.cpp:665-718 (+ Enzyme adjoint :270-279)
//   K3 tv / adam / fill : ref kernel2D_TVGrad :737-749, optimizerStep example_tracer.h:128-206, zeroGrad :104-109
//   K4 upsample / boxes : ref UpsampleGrid :281-323, UpsampleOctree :325-447, RebuildOctree :449-489
//
// One thread per ray; a warp covers an 8x4 pixel tile so that its 32 rays walk neighbouring voxels.
// Built with -fmad=false (see lrf_device.cuh): decision-critical arithmetic is IEEE op-for-op the
// reference's; explicit fmaf() only on the colour path (RGB tolerance 1e-4).
#include <algorithm>
#include "lrf_kernels.cuh"
#include "lrf_fast.cuh"
#include "lrf_lean.cuh"
#include "lrf_wray.cuh"
#include "lrf_staged.cuh"
#include "lrf_tv.cuh"

namespace lrf {

namespace {

constexpr int kBlock = LRF_KBLOCK;
static_assert(kBlock == kFastBlock, "FastMarch's shared-memory stride assumes the forward kernels' block size");

struct RayCtx
{
  float o[3], d[3], idir[3];
  float Y[9];
};

__device__ __forceinline__ void ray_ctx_init(RayCtx &R, const float *o, const float *d, bool grey)
{
#pragma unroll
  for (int a = 0; a < 3; a++) {
    R.o[a] = o[a];
    R.d[a] = d[a];
    R.idir[a] = 1.0f / d[a];  // ref: example_tracer.cpp:20-22 (recomputed per box there; same value)
  }
  if (!grey) sh_eval_2(R.d, R.Y);
}

struct SampleGeom
{
  float t, delta, lf[3];
  uint32_t base;  // node index of the near corner
};

// ref: RayGridIntersection example_tracer.cpp:134-153 — geometry of one sample, bit-exact
__device__ __forceinline__ void sample_geom(const SceneDev &S, const RayCtx &R, int x, int y, int z, float &tPrev, SampleGeom &g)
{
  const int c[3] = {x, y, z};
  float bmin[3], bmax[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    bmin[a] = (float)c[a] * S.inv_gs_f;        // == (float)x / (float)gs for gs = 2^k  (example_tracer.h:80)
    bmax[a] = (float)(c[a] + 1) * S.inv_gs_f;
  }
  // ref: RayBoxIntersection example_tracer.cpp:18-43
  float lo = R.idir[0] * (bmin[0] - R.o[0]), hi = R.idir[0] * (bmax[0] - R.o[0]);
  float tmin = lm_min(lo, hi), tmax = lm_max(lo, hi);
  lo = R.idir[1] * (bmin[1] - R.o[1]); hi = R.idir[1] * (bmax[1] - R.o[1]);
  tmin = lm_max(tmin, lm_min(lo, hi)); tmax = lm_min(tmax, lm_max(lo, hi));
  lo = R.idir[2] * (bmin[2] - R.o[2]); hi = R.idir[2] * (bmax[2] - R.o[2]);
  tmin = lm_max(tmin, lm_min(lo, hi)); tmax = lm_min(tmax, lm_max(lo, hi));
  if (tPrev < 0.0f) tPrev = tmin;                            // :139
  const float t = (tmin + tmax) * 0.5f;                      // :142  (/2.0f is exact either way)
  float actual = t - tPrev;                                  // :144
  if (actual > S.step_gate_f)                                // :145  actual > sqrt(3)*step, compared in double there
    actual = (float)((double)actual - (double)((int)((double)actual / S.step) - 1) * S.step);  // :146
  g.t = t;
  g.delta = actual;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float p = R.o[a] + t * R.d[a];                     // :148
    g.lf[a] = (p - bmin[a]) * S.gs_f;                        // :150  division by (bmax-bmin) = 2^-k is this exact scaling
  }
  g.base = (uint32_t)x + (uint32_t)y * S.gs + (uint32_t)z * S.gs * S.gs;   // :152-153 near = (x,y,z), far = near + 1
  tPrev = t;                                                 // :190
}

// trilinear density, nested lerps in the reference's order (:155-167 via lerpCell :72-80), bit-exact
__device__ __forceinline__ float density_at(const SceneDev &S, const SampleGeom &g)
{
  const float *p = S.dens + g.base;
  const uint32_t sy = S.gs, sz = (uint32_t)S.gs * S.gs;
  const float d000 = __ldg(p), d100 = __ldg(p + 1), d010 = __ldg(p + sy), d110 = __ldg(p + sy + 1);
  const float d001 = __ldg(p + sz), d101 = __ldg(p + sz + 1), d011 = __ldg(p + sz + sy), d111 = __ldg(p + sz + sy + 1);
  const float xy00 = lm_lerp(d000, d100, g.lf[0]), xy10 = lm_lerp(d010, d110, g.lf[0]);
  const float xy01 = lm_lerp(d001, d101, g.lf[0]), xy11 = lm_lerp(d011, d111, g.lf[0]);
  const float xyz0 = lm_lerp(xy00, xy10, g.lf[1]), xyz1 = lm_lerp(xy01, xy11, g.lf[1]);
  return lm_lerp(xyz0, xyz1, g.lf[2]);
}

// ------------------------------------------------------------------------------------------------
// Register window over the 2x2x2 nodes of the current voxel.
//
// Consecutive samples of a ray sit in face-adjacent voxels and share 4 of their 8 nodes, so re-gathering
// all 8 x 52 B per sample doubles the L1 traffic — and the first ncu capture of this kernel shows the L1
// data pipe (l1tex__data_pipe_lsu_wavefronts) at 88 % with DRAM at 1 %: the gather, not HBM, is the bound.
// The window keeps, per node, the density and the colour contracted with this ray's SH basis
// {R, G, B, q = sum_j sh_j Y_j(d)} in registers.  A node lives in the slot given by the PARITY of its
// coordinates (slot = x&1 | (y&1)<<1 | (z&1)<<2): when the ray moves to a face-adjacent voxel the 4 nodes
// that stay keep their slots and the 4 new ones have exactly the parities of the 4 that leave, so nothing
// is ever moved between registers — only (predicated) loads into the slots whose coordinates changed, and
// the interpolation weights are permuted instead of the values.  Any transition (first sample, the 3/7-voxel
// tie events, gaps of a sparse box set) is handled by the same compare-and-reload.
struct Window
{
  float d[8];
  float r[8], g[8], b[8], q[8];
  int cd[3][2];   // node coordinate currently held by parity 0 / 1 along each axis (density part)
  int cc[3][2];   // same for the colour part (loaded lazily: only samples with a non-zero weight need it)
  __device__ __forceinline__ void reset()
  {
#pragma unroll
    for (int a = 0; a < 3; a++) { cd[a][0] = cd[a][1] = cc[a][0] = cc[a][1] = -4; }
  }
};

// coordinates of the voxel's two nodes along one axis, by parity: parity-0 slot holds v + (v&1), parity-1 slot v + 1 - (v&1)
__device__ __forceinline__ void window_density(const SceneDev &S, Window &W, int x, int y, int z)
{
  const int v[3] = {x, y, z};
  int req[3][2];
  bool chg[3][2];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    req[a][0] = v[a] + (v[a] & 1);
    req[a][1] = v[a] + 1 - (v[a] & 1);
    chg[a][0] = W.cd[a][0] != req[a][0];
    chg[a][1] = W.cd[a][1] != req[a][1];
    W.cd[a][0] = req[a][0];
    W.cd[a][1] = req[a][1];
  }
  const uint32_t gs = S.gs, gs2 = (uint32_t)S.gs * S.gs;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const int sx = s & 1, sy = (s >> 1) & 1, sz = s >> 2;
    if (chg[0][sx] | chg[1][sy] | chg[2][sz])
      W.d[s] = __ldg(S.dens + (uint32_t)req[0][sx] + (uint32_t)req[1][sy] * gs + (uint32_t)req[2][sz] * gs2);
  }
}

// trilinear density from the window: nested lerps in the reference's order (:155-167 via lerpCell :72-80), bit-exact.
// near/far along an axis = parity slot (v&1) / the other one.
__device__ __forceinline__ float window_density_lerp(const Window &W, const SampleGeom &g, int x, int y, int z)
{
  const bool ox = x & 1, oy = y & 1, oz = z & 1;
  float A[4];  // x-lerped, indexed by (sy, sz) parity
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float p0 = W.d[2 * j], p1 = W.d[2 * j + 1];
    A[j] = lm_lerp(ox ? p1 : p0, ox ? p0 : p1, g.lf[0]);
  }
  float B[2];  // y-lerped, indexed by sz parity
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const float p0 = A[2 * j], p1 = A[2 * j + 1];
    B[j] = lm_lerp(oy ? p1 : p0, oy ? p0 : p1, g.lf[1]);
  }
  return lm_lerp(oz ? B[1] : B[0], oz ? B[0] : B[1], g.lf[2]);
}

// interpolation weight of each parity slot (explicit-FMA colour path; also the scatter weights of the backward)
__device__ __forceinline__ void window_weights(const SampleGeom &g, int x, int y, int z, float *w)
{
  const float fx = g.lf[0], fy = g.lf[1], fz = g.lf[2];
  const float gx = 1.0f - fx, gy = 1.0f - fy, gz = 1.0f - fz;
  const bool ox = x & 1, oy = y & 1, oz = z & 1;
  const float wx0 = ox ? fx : gx, wx1 = ox ? gx : fx;
  const float wy0 = oy ? fy : gy, wy1 = oy ? gy : fy;
  const float wz0 = oz ? fz : gz, wz1 = oz ? gz : fz;
  const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
  w[0] = w00 * wz0; w[1] = w10 * wz0; w[2] = w01 * wz0; w[3] = w11 * wz0;
  w[4] = w00 * wz1; w[5] = w10 * wz1; w[6] = w01 * wz1; w[7] = w11 * wz1;
}

// bring the colour part of the window to voxel (x,y,z): load the nodes whose coordinates changed, contracting the 9 SH
// coefficients with this ray's basis on the fly
__device__ __forceinline__ void window_colour(const SceneDev &S, Window &W, const float *Y, int x, int y, int z)
{
  const int v[3] = {x, y, z};
  int req[3][2];
  bool chg[3][2];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    req[a][0] = v[a] + (v[a] & 1);
    req[a][1] = v[a] + 1 - (v[a] & 1);
    chg[a][0] = W.cc[a][0] != req[a][0];
    chg[a][1] = W.cc[a][1] != req[a][1];
    W.cc[a][0] = req[a][0];
    W.cc[a][1] = req[a][1];
  }
  const uint32_t gs = S.gs, gs2 = (uint32_t)S.gs * S.gs;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const int sx = s & 1, sy = (s >> 1) & 1, sz = s >> 2;
    if (chg[0][sx] | chg[1][sy] | chg[2][sz]) {
      const uint32_t idx = (uint32_t)req[0][sx] + (uint32_t)req[1][sy] * gs + (uint32_t)req[2][sz] * gs2;
      const float4 *p = S.col + (size_t)idx * kColF4;
      const float4 a = __ldg(p), b = __ldg(p + 1), e = __ldg(p + 2);
      float qq = a.w * Y[0];
      qq = fmaf(b.x, Y[1], qq); qq = fmaf(b.y, Y[2], qq); qq = fmaf(b.z, Y[3], qq); qq = fmaf(b.w, Y[4], qq);
      qq = fmaf(e.x, Y[5], qq); qq = fmaf(e.y, Y[6], qq); qq = fmaf(e.z, Y[7], qq); qq = fmaf(e.w, Y[8], qq);
      W.r[s] = a.x; W.g[s] = a.y; W.b[s] = a.z; W.q[s] = qq;
    }
  }
}

// weighted sums over the window + clamps (ref :181-184): rgbi = interpolated RGB, raw = interpolated sh . Y
__device__ __forceinline__ void window_colour_eval(const Window &W, const float *w, float *rgbi, float &raw, float &s, float *c)
{
  float R = 0.0f, G = 0.0f, B = 0.0f, Q = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    R = fmaf(w[k], W.r[k], R); G = fmaf(w[k], W.g[k], G); B = fmaf(w[k], W.b[k], B); Q = fmaf(w[k], W.q[k], Q);
  }
  rgbi[0] = R; rgbi[1] = G; rgbi[2] = B;
  raw = Q;
  s = lm_clamp(Q, 0.0f, 1.0f);
  c[0] = lm_clamp(R * s, 0.0f, 1.0f); c[1] = lm_clamp(G * s, 0.0f, 1.0f); c[2] = lm_clamp(B * s, 0.0f, 1.0f);
}

// ------------------------------------------------------------------------------------------------
// forward visitor.  MODE 0: render (stop at the reference's throughput break);  MODE 1: parity taps
// (keeps marching after the break to count every hit, records voxels / cauchy / touched nodes)
template <int MODE> struct ShadeVisitor
{
  const SceneDev &S;
  const uint64_t *etab;
  RayCtx R;
  bool grey;
  float thr = 1.0f, tPrev = -1.0f, cauchy = 0.0f;
  float col[3] = {0.0f, 0.0f, 0.0f};
  uint32_t n_hit = 0, n_shaded = 0;
  bool alive = true;
  // taps
  int32_t *vox = nullptr;
  uint32_t keep = 0;
  uint8_t *touched = nullptr;
  Window W;

  __device__ ShadeVisitor(const SceneDev &s, const uint64_t *t, bool g) : S(s), etab(t), grey(g) { W.reset(); }

  __device__ __forceinline__ bool sample(int x, int y, int z)
  {
    if (MODE == 1) {
      if (vox && n_hit < keep) { vox[n_hit * 3] = x; vox[n_hit * 3 + 1] = y; vox[n_hit * 3 + 2] = z; }
      n_hit++;
      if (!alive) return true;
    } else n_hit = 1;
    SampleGeom g;
    sample_geom(S, R, x, y, z, tPrev, g);
    window_density(S, W, x, y, z);
    float sigma = window_density_lerp(W, g, x, y, z);
    if (sigma < 0.0f) sigma = 0.0f;                          // :170-171
    if (MODE == 1) {
      cauchy += logf(1.0f + 2.0f * sigma * sigma);           // :173 (value only feeds the loss; tolerance-compared)
      if (touched) {
        const uint32_t sy = S.gs, sz = (uint32_t)S.gs * S.gs;
#pragma unroll
        for (int c = 0; c < 8; c++) touched[g.base + (c & 1) + ((c >> 1) & 1) * sy + ((c >> 2) & 1) * sz] = 1;
      }
    }
    const float tr = expf_libm(-sigma * g.delta, etab);      // :175
    const float wgt = thr * (1.0f - tr);                     // :187
    if (wgt != 0.0f) {                                       // wgt == 0 adds exactly 0 to a finite colour: skip the gather
      if (grey) { col[0] = col[0] + wgt; col[1] = col[1] + wgt; col[2] = col[2] + wgt; }
      else {
        float w[8], rgbi[3], raw, s, c[3];
        window_colour(S, W, R.Y, x, y, z);
        window_weights(g, x, y, z, w);
        window_colour_eval(W, w, rgbi, raw, s, c);
        col[0] = col[0] + wgt * c[0]; col[1] = col[1] + wgt * c[1]; col[2] = col[2] + wgt * c[2];
      }
    }
    thr *= tr;                                               // :189
    n_shaded++;
    if (thr < 0.1f) {                                        // :131  (float < 0.1 as double  <=>  float < 0.1f)
      alive = false;
      return MODE == 1;
    }
    return true;
  }
};

__device__ __forceinline__ void load_exp_table(uint64_t *tab)
{
  if (threadIdx.x < 32) tab[threadIdx.x] = kExp2fTab[threadIdx.x];
  __syncthreads();
}

// pixel of this thread: a warp covers an 8x4 pixel tile, each quarter-warp a 4x2 block of it.  The L1 data pipe serves
// a 128-bit warp load one quarter-warp at a time and spends one wavefront per distinct 128-byte line in it, so the
// more compact the quarter's footprint on the grid, the fewer wavefronts a gather costs (ncu: 8.3 per LDG.128 with
// 8x1 quarters; the pipe was 88 % busy).
#ifndef LRF_TILE2D
#define LRF_TILE2D 1   // 1 = a 128-thread block covers 16 x 8 pixels (2 x 2 warp tiles) instead of a 32 x 4 strip: its warps share more nodes in L1
                       // (sparse 512^3 render 460 -> 476 Mrays/s, whole-image backward 12.07 -> 11.99 ms; r2r variants34.log)
#endif
__device__ __forceinline__ bool tile_pixel(uint32_t w, uint32_t h, uint32_t &px, uint32_t &py, uint32_t bid = blockIdx.x)
{
  const uint32_t lane = threadIdx.x & 31, q = lane >> 3, sub = lane & 7;
#if LRF_TILE2D
  static_assert(kBlock == 128, "2 x 2 warp tiles per block");
  const uint32_t bx_n = (w + 15) / 16, wq = threadIdx.x >> 5;
  px = (bid % bx_n) * 16 + (wq & 1) * 8 + (q & 1) * 4 + (sub & 3);
  py = (bid / bx_n) * 8 + (wq >> 1) * 4 + (q >> 1) * 2 + (sub >> 2);
#else
  const uint32_t tiles_x = (w + 7) / 8;
  const uint32_t tile = bid * (kBlock / 32) + (threadIdx.x >> 5);
  px = (tile % tiles_x) * 8 + (q & 1) * 4 + (sub & 3);
  py = (tile / tiles_x) * 4 + (q >> 1) * 2 + (sub >> 2);
#endif
  return px < w && py < h;
}
inline uint32_t tile_grid(uint32_t w, uint32_t h)
{
#if LRF_TILE2D
  return ((w + 15) / 16) * ((h + 7) / 8);
#else
  const uint32_t tiles = ((w + 7) / 8) * ((h + 3) / 4);
  return (tiles + (kBlock / 32) - 1) / (kBlock / 32);
#endif
}

#ifndef LRF_RENDER_MINB
#define LRF_RENDER_MINB 4   // resident blocks per SM the register allocation is sized for (tools/variants.py)
#endif
#ifndef LRF_SPARSE_MINB
#define LRF_SPARSE_MINB 3   // the same for the sparse instantiations: 168 registers, no spills — with one traversal event per trip (LRF_SEARCH_TRIPS = 1)
                            // the 512^3 sphere scene renders at 463 Mrays/s in the batch launch against 393 at 4 blocks / 4 events (r2q, variants31.log)
#endif

// Forward render, one thread per pixel.  Rays that cannot take the fast path (a zero / non-finite reciprocal direction
// component) are appended to a fallback list and rendered by render_fallback_kernel with the generic march.
//
// Tried and dropped (profiles/r1f_variants.txt box 12): a persistent grid of 148 x 4 blocks pulling 8x4-pixel tiles in Morton
// order from per-SM queues (%smid-affine, so that the warps sharing an L1 work on adjacent pixels; with stealing), from 8
// queues and from one global queue, per block and per warp: 2.40-2.50 ms per view against 2.30 ms for the plain grid.  The
// hardware's in-order block dispatch keeps the whole GPU on one compact band of the image, and that L2 locality is worth
// more than the L1 locality or the finer load balance.
// executed by all 32 lanes of a warp together (valid = the lane has a pixel): the sparse march votes among the lanes that run it
template <bool SPARSE, bool GREY, bool BAKED = false>
__device__ __forceinline__ void render_pixel(const SceneDev &S, const uint64_t *etab, float *ysm, float4 *csm, uint32_t *__restrict__ out, uint32_t w,
                                             uint32_t h, uint32_t px, uint32_t py, bool valid, uint32_t *__restrict__ fb_count,
                                             uint32_t *__restrict__ fb_list, const float *view_inv, uint32_t frame0)
{
  float o[3], d[3];
  gen_ray(view_inv, S.proj_inv, valid ? px : 0u, valid ? py : 0u, w, h, o, d);
  const bool fast = fast_path_ok(d);
  const uint32_t mask = __ballot_sync(0xffffffffu, valid && fast);
  if (!valid) return;
  if (!fast) {
    fb_list[atomicAdd(fb_count, 1u)] = frame0 + py * w + px;
    return;
  }
  FastMarch<0, SPARSE, GREY, BAKED> fm(S, etab, ysm + threadIdx.x, csm + threadIdx.x);
  fm.run(o, d, mask);
  uint32_t pix = 0;  // ref :645,660: rays with an empty hit list pack float4(0)
  if (fm.n_hit)
    pix = pack_color(lm_clamp(fm.col[0], 0.0f, 1.0f), lm_clamp(fm.col[1], 0.0f, 1.0f), lm_clamp(fm.col[2], 0.0f, 1.0f), 1.0f);  // :193,:653
  out[(size_t)frame0 + (size_t)py * w + px] = pix;
}

// Rows [y0, y0 + hb) of the w x h frame (the whole frame: y0 = 0, hb = h).  Row bands are how one frame is split over several GPUs
// (latency mode; the reference splits the rows of a frame over its OpenMP threads, ref example_tracer.cpp:630).
template <bool SPARSE, bool GREY, bool BAKED = false>
__global__ void __launch_bounds__(kBlock, (SPARSE ? LRF_SPARSE_MINB : LRF_RENDER_MINB) * 128 / kBlock) render_image_kernel(const __grid_constant__ SceneDev S, uint32_t *__restrict__ out, uint32_t w, uint32_t h,
                                                                 uint32_t *__restrict__ fb_count, uint32_t *__restrict__ fb_list, uint32_t y0, uint32_t hb,
                                                                 const float *__restrict__ views = nullptr, uint32_t blocks_per_view = 0)
{
  __shared__ uint64_t etab[32];
  __shared__ float ysm[9 * kFastBlock];
  __shared__ float4 csm[8 * kFastBlock];
  load_exp_table(etab);
  // views != nullptr: a batch of views in one launch (see render_lean_kernel<.., BATCH>): block b = block b % blocks_per_view of view b / blocks_per_view
  uint32_t px, py, bid = blockIdx.x, frame0 = 0;
  const float *view_inv = S.view_inv;
  if (views) {
    const uint32_t v = bid / blocks_per_view;
    bid -= v * blocks_per_view;
    view_inv = views + 16 * v;
    frame0 = v * w * h;
  }
  const bool valid = tile_pixel(w, hb, px, py, bid);
  py += y0;
  render_pixel<SPARSE, GREY, BAKED>(S, etab, ysm, csm, out, w, h, px, py, valid, fb_count, fb_list, view_inv, frame0);
}

// Dense colour frames: the register-window march of lrf_lean.cuh (same decisions as FastMarch, bit for bit; fewer instructions and
// L1 wavefronts per sample).  tap_n / tap_thr (TAPS): per-pixel shaded-sample count and final transmittance for the parity tests
// (0xffffffff / NaN where the pixel went to the fallback list).
#ifndef LRF_LEAN_TILE2D
#define LRF_LEAN_TILE2D 1   // blocks of 16 x 8 pixels (2 x 2 warp tiles); 0 = 32 x 4 strips (measured: +0.5 % with 1)
#endif
#ifndef LRF_LEAN_BLOCK
#define LRF_LEAN_BLOCK 32    // threads per block: 128 = 16 x 8 pixels, 64 = 16 x 4, 32 = 8 x 4 (one warp tile: no warp waits for its block's slowest; +0.7 %)
#endif
#ifndef LRF_LEAN_WARPS
#define LRF_LEAN_WARPS (LRF_LEAN_MINB * 4)   // resident warps per SM the register allocation is sized for (12 -> 168 registers)
#endif
#ifndef LRF_LEAN_SUPER
#define LRF_LEAN_SUPER 0     // 1 = a block of LRF_LEAN_BLOCK threads owns a 32 x 32-pixel super-tile and its warps pull the 32 warp tiles of it from a shared
                             // counter (neighbouring tiles stay on one SM: L1 sharing of the 16 x 8 blocks, with the balance of one-warp blocks)
#endif
constexpr uint32_t kLeanBW = LRF_LEAN_SUPER ? 32 : (LRF_LEAN_BLOCK >= 64 ? 16 : 8), kLeanBH = LRF_LEAN_SUPER ? 32 : (LRF_LEAN_BLOCK >= 128 ? 8 : 4);
// BATCH: one launch renders n views (lrf_render_views_dev_async): block b works on tile b % tiles of view b / tiles, reads that view's
// matrix from `views` (device, n x 16 floats) and writes frame b / tiles of `out` (frames w * h pixels apart).  Per-view launches leave
// the SMs idle while the last blocks of a view finish (ncu: SMs active 95 % of the launch, 11.3 of 12 warps resident on average); in
// one launch the next view's blocks move in as the previous view's drain, in the same order the per-view launches would run them.
template <bool TAPS, bool GATE, bool BATCH = false>
__global__ void __launch_bounds__(LRF_LEAN_BLOCK, LRF_LEAN_WARPS * 32 / LRF_LEAN_BLOCK) render_lean_kernel(const __grid_constant__ SceneDev S, uint32_t *__restrict__ out, uint32_t w, uint32_t h,
                                                                            uint32_t *__restrict__ fb_count, uint32_t *__restrict__ fb_list, uint32_t y0, uint32_t hb,
                                                                            uint32_t *__restrict__ tap_n, float *__restrict__ tap_thr,
                                                                            const float *__restrict__ views = nullptr, uint32_t tiles = 0)
{
  __shared__ uint64_t etab[32];
#if LRF_LEAN_SUPER
  __shared__ uint32_t next_tile;
  if (threadIdx.x == 0) next_tile = LRF_LEAN_BLOCK / 32;   // the first tiles are the warps' own
#endif
  load_exp_table(etab);
  uint32_t px, py, bid = blockIdx.x, frame0 = 0;
  const float *view_inv = S.view_inv;
  if (BATCH) {
    const uint32_t v = bid / tiles;
    bid -= v * tiles;
    view_inv = views + 16 * v;
    frame0 = v * w * h;                  // < 2^32: checked by the host
  }
  auto render_one = [&](uint32_t px, uint32_t py) {
    float o[3], d[3];
    gen_ray(view_inv, S.proj_inv, px, py, w, h, o, d);
    const size_t pixel = (size_t)frame0 + (size_t)py * w + px;
    if (!fast_path_ok(d)) {
      fb_list[atomicAdd(fb_count, 1u)] = (uint32_t)pixel;
      if (TAPS) { tap_n[pixel] = 0xffffffffu; tap_thr[pixel] = __int_as_float(0x7fc00000); }
      return;
    }
    LeanMarch<TAPS, GATE> lm(S, etab);
    lm.run(o, d);
    uint32_t pix = 0;  // ref :645,660: rays with an empty hit list pack float4(0)
    if (lm.any_hit())
      pix = pack_color(lm_clamp(lm.col[0], 0.0f, 1.0f), lm_clamp(lm.col[1], 0.0f, 1.0f), lm_clamp(lm.col[2], 0.0f, 1.0f), 1.0f);  // :193,:653
    out[pixel] = pix;
    if (TAPS) { tap_n[pixel] = lm.n_shaded; tap_thr[pixel] = lm.thr; }
  };
#if LRF_LEAN_SUPER
  {
    static_assert(LRF_LEAN_TILE2D, "super-tiles are a 2-D tiling");
    const uint32_t bx_n = (w + 31) / 32, sx = bid % bx_n, sy = bid / bx_n, lane = threadIdx.x & 31, q = lane >> 3, sub = lane & 7;
    for (uint32_t tidx = threadIdx.x >> 5; tidx < 32;) {
      // tile order: 2 x 2 groups (16 x 8 pixels), groups row by row — the warps of a block work on neighbouring tiles
      const uint32_t g = tidx >> 2, s4 = tidx & 3, tx = (g & 1) * 2 + (s4 & 1), ty = (g >> 1) * 2 + (s4 >> 1);
      px = sx * 32 + tx * 8 + (q & 1) * 4 + (sub & 3);
      py = sy * 32 + ty * 4 + (q >> 1) * 2 + (sub >> 2);
      if (px < w && py < hb) render_one(px, py + y0);
      uint32_t t = 0;
      if (lane == 0) t = atomicAdd(&next_tile, 1u);
      tidx = __shfl_sync(0xffffffffu, t, 0);
    }
  }
#elif LRF_LEAN_TILE2D
  {
    // a block covers 16 x 8 pixels (2 x 2 warp tiles of 8 x 4) instead of a 32 x 4 strip: its four warps share more nodes in L1
    const uint32_t bx_n = (w + kLeanBW - 1) / kLeanBW;
    const uint32_t bx = bid % bx_n, by = bid / bx_n, wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = lane >> 3, sub = lane & 7;
    px = bx * kLeanBW + (wq & 1) * 8 + (q & 1) * 4 + (sub & 3);
    py = by * kLeanBH + (wq >> 1) * 4 + (q >> 1) * 2 + (sub >> 2);
    if (px >= w || py >= hb) return;
    render_one(px, py + y0);
  }
#else
  static_assert(!BATCH, "the batch launch needs the 2-D block tiling");
  if (!tile_pixel(w, hb, px, py)) return;
  render_one(px, py + y0);
#endif
}

// views != nullptr: the list holds frame * w * h + pixel of a batch of views (render_lean_kernel<.., BATCH>)
__global__ void __launch_bounds__(kBlock) render_fallback_kernel(const __grid_constant__ SceneDev S, uint32_t *__restrict__ out, uint32_t w, uint32_t h, int grey,
                                                                 const uint32_t *__restrict__ fb_count, const uint32_t *__restrict__ fb_list,
                                                                 const float *__restrict__ views = nullptr)
{
  __shared__ uint64_t etab[32];
  load_exp_table(etab);
  const uint32_t n = *fb_count;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    const uint32_t pixel = fb_list[i], frame = pixel / (w * h), in_frame = pixel - frame * (w * h), px = in_frame % w, py = in_frame / w;
    float o[3], d[3];
    gen_ray(views ? views + 16 * frame : S.view_inv, S.proj_inv, px, py, w, h, o, d);
    ShadeVisitor<0> vis(S, etab, grey != 0);
    ray_ctx_init(vis.R, o, d, grey != 0);
    march(S, o, d, vis);
    uint32_t pix = 0;
    if (vis.n_hit)
      pix = pack_color(lm_clamp(vis.col[0], 0.0f, 1.0f), lm_clamp(vis.col[1], 0.0f, 1.0f), lm_clamp(vis.col[2], 0.0f, 1.0f), 1.0f);
    out[pixel] = pix;
  }
}

// Exact render of the rays the staged kernel (lrf_staged.cuh) did not decide: the streamlined exact march where it applies, the
// generic fp64 march for the rest (axis-aligned rays).  Dense sets only.
__global__ void __launch_bounds__(kBlock) render_list_kernel(const __grid_constant__ SceneDev S, uint32_t *__restrict__ out, uint32_t w, uint32_t h,
                                                             const uint32_t *__restrict__ fb_count, const uint32_t *__restrict__ fb_list)
{
  __shared__ uint64_t etab[32];
  __shared__ float ysm[9 * kFastBlock];
  __shared__ float4 csm[8 * kFastBlock];
  load_exp_table(etab);
  const uint32_t n = *fb_count;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    const uint32_t pixel = fb_list[i], px = pixel % w, py = pixel / w;
    float o[3], d[3];
    gen_ray(S.view_inv, S.proj_inv, px, py, w, h, o, d);
    uint32_t pix = 0;
    if (fast_path_ok(d)) {
      FastMarch<0, false, false> fm(S, etab, ysm + threadIdx.x, csm + threadIdx.x);
      fm.run(o, d);
      if (fm.n_hit) pix = pack_color(lm_clamp(fm.col[0], 0.0f, 1.0f), lm_clamp(fm.col[1], 0.0f, 1.0f), lm_clamp(fm.col[2], 0.0f, 1.0f), 1.0f);
    } else {
      ShadeVisitor<0> vis(S, etab, false);
      ray_ctx_init(vis.R, o, d, false);
      march(S, o, d, vis);
      if (vis.n_hit) pix = pack_color(lm_clamp(vis.col[0], 0.0f, 1.0f), lm_clamp(vis.col[1], 0.0f, 1.0f), lm_clamp(vis.col[2], 0.0f, 1.0f), 1.0f);
    }
    out[pixel] = pix;
  }
}

// parity taps.  force_generic = 1 sends every ray down the generic march (tests compare both paths)
template <bool IMAGE>
__global__ void __launch_bounds__(kBlock) debug_kernel(const __grid_constant__ SceneDev S, uint32_t w, uint32_t h, size_t n_rays, const float *__restrict__ ro,
                                                       const float *__restrict__ rd, int grey, DebugTaps taps, int force_generic)
{
  __shared__ uint64_t etab[32];
  __shared__ float ysm[9 * kFastBlock];
  __shared__ float4 csm[8 * kFastBlock];
  load_exp_table(etab);
  size_t ray;
  float o[3], d[3];
  if (IMAGE) {
    uint32_t px, py;
    if (!tile_pixel(w, h, px, py)) return;
    ray = (size_t)py * w + px;
    gen_ray(S.view_inv, S.proj_inv, px, py, w, h, o, d);
  } else {
    ray = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (ray >= n_rays) return;
#pragma unroll
    for (int a = 0; a < 3; a++) { o[a] = ro[ray * 3 + a]; d[a] = rd[ray * 3 + a]; }
  }
  int32_t *vox = taps.vox_xyz ? taps.vox_xyz + ray * taps.keep * 3 : nullptr;
  if (vox)
    for (uint32_t i = 0; i < taps.keep * 3; i++) vox[i] = -1;
  uint32_t n_hit, n_shaded;
  float col[3], thr, cauchy;
  if (!force_generic && fast_path_ok(d)) {
    auto run_fast = [&](auto &fm) {
      fm.keep = taps.keep;
      fm.vox = vox;
      fm.touched = taps.touched;
      fm.run(o, d);
      n_hit = fm.n_hit; n_shaded = fm.n_shaded; thr = fm.thr; cauchy = fm.cauchy;
      col[0] = fm.col[0]; col[1] = fm.col[1]; col[2] = fm.col[2];
    };
    if (grey) { FastMarch<1, true, true> fm(S, etab, ysm + threadIdx.x, csm + threadIdx.x); run_fast(fm); }
    else { FastMarch<1, true, false> fm(S, etab, ysm + threadIdx.x, csm + threadIdx.x); run_fast(fm); }
  } else {
    ShadeVisitor<1> vis(S, etab, grey != 0);
    ray_ctx_init(vis.R, o, d, grey != 0);
    vis.keep = taps.keep;
    vis.vox = vox;
    vis.touched = taps.touched;
    march(S, o, d, vis);
    n_hit = vis.n_hit; n_shaded = vis.n_shaded; thr = vis.thr; cauchy = vis.cauchy;
    col[0] = vis.col[0]; col[1] = vis.col[1]; col[2] = vis.col[2];
  }
  if (taps.n_hit) taps.n_hit[ray] = n_hit;
  if (taps.n_shaded) taps.n_shaded[ray] = n_shaded;
  if (taps.rgbtc) {
    float *r = taps.rgbtc + ray * 5;
    if (n_hit) {
      r[0] = lm_clamp(col[0], 0.0f, 1.0f); r[1] = lm_clamp(col[1], 0.0f, 1.0f); r[2] = lm_clamp(col[2], 0.0f, 1.0f);
      r[3] = thr; r[4] = cauchy;
    } else { r[0] = r[1] = r[2] = r[3] = r[4] = 0.0f; }
  }
}

// ------------------------------------------------------------------------------------------------
// backward.  Pass A = the forward (colour sums kept per channel in fp64 so that pass B can form the
// exact suffix  sum_{m>k} T_m (1-tr_m) c_m  as  total - prefix  without cancellation noise);
// pass B = forward again, emitting the adjoint of every sample in FORWARD order (no tape in HBM):
//   a_tr_k   = -T_k (c_k . aC) + [ (suffix_k . aC) + aT_n T_n ] / tr_k           (k < n-1; tr_k >= 0.1 there)
//   a_tr_n-1 = -T_k (c_k . aC) + aT_n T_k
// which is the closed form of the reverse sweep  aT <- tr aT + (1-tr) (c . aC)  of SURVEY 8 a11.
struct FwdAVisitor
{
  const SceneDev &S;
  const uint64_t *etab;
  RayCtx R;
  bool grey;
  float thr = 1.0f, tPrev = -1.0f;
  double csum[3] = {0.0, 0.0, 0.0};
  float col[3] = {0.0f, 0.0f, 0.0f};
  uint32_t n_shaded = 0;
  Window W;
  __device__ FwdAVisitor(const SceneDev &s, const uint64_t *t, bool g) : S(s), etab(t), grey(g) { W.reset(); }
  __device__ __forceinline__ bool sample(int x, int y, int z)
  {
    SampleGeom g;
    sample_geom(S, R, x, y, z, tPrev, g);
    window_density(S, W, x, y, z);
    float sigma = window_density_lerp(W, g, x, y, z);
    if (sigma < 0.0f) sigma = 0.0f;
    const float tr = expf_libm(-sigma * g.delta, etab);
    const float wgt = thr * (1.0f - tr);
    if (wgt != 0.0f) {
      float c[3] = {1.0f, 1.0f, 1.0f};
      if (!grey) {
        float w[8], rgbi[3], raw, s;
        window_colour(S, W, R.Y, x, y, z);
        window_weights(g, x, y, z, w);
        window_colour_eval(W, w, rgbi, raw, s, c);
      }
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const float term = wgt * c[ch];
        col[ch] = col[ch] + term;
        csum[ch] += (double)term;
      }
    }
    thr *= tr;
    n_shaded++;
    return !(thr < 0.1f);
  }
};

__device__ __forceinline__ void red_add(float *p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add4(float4 *p, float a, float b, float c, float d)
{
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct BwdVisitor
{
  const SceneDev &S;
  const uint64_t *etab;
  RayCtx R;
  bool grey;
  float thr = 1.0f, tPrev = -1.0f;
  double cpre[3] = {0.0, 0.0, 0.0};
  // from pass A
  double ctot[3];
  float aC[3], aTn, Tn;
  uint32_t n_total, k = 0;
  Window W;
  // Gradient window.  Consecutive samples of a ray share 4 of their 8 nodes, and for one ray the 13 adjoints of a node are
  // 5 numbers — density, R, G, B and the adjoint of q = sh . Y(d) (the 9 SH adjoints are that one number times the ray's
  // basis).  They are accumulated per parity slot and leave for HBM only when the slot's node changes (or the ray ends):
  // 4 nodes x 4 red.global.add per step instead of 8 x 4, and 5 FMAs per corner instead of 13 multiplies.
  float ad[8], ar[8], ag[8], ab[8], aq[8];
  int ca[3][2];     // node coordinates the slots currently accumulate for (-4: empty)
  __device__ BwdVisitor(const SceneDev &s, const uint64_t *t, bool g) : S(s), etab(t), grey(g)
  {
    W.reset();
#pragma unroll
    for (int a = 0; a < 3; a++) ca[a][0] = ca[a][1] = -4;
#pragma unroll
    for (int i = 0; i < 8; i++) ad[i] = ar[i] = ag[i] = ab[i] = aq[i] = 0.0f;
  }
  template <int SLOT> __device__ __forceinline__ void flush_slot()
  {
    const int sx = SLOT & 1, sy = (SLOT >> 1) & 1, sz = SLOT >> 2;
    const uint32_t node = (uint32_t)ca[0][sx] + (uint32_t)ca[1][sy] * S.gs + (uint32_t)ca[2][sz] * S.gs * S.gs;
    if (ad[SLOT] != 0.0f) red_add(S.gdens + node, ad[SLOT]);
    if (!grey && (aq[SLOT] != 0.0f || ar[SLOT] != 0.0f || ag[SLOT] != 0.0f || ab[SLOT] != 0.0f)) {
      float4 *q = S.gcol + (size_t)node * kColF4;
      const float v = aq[SLOT];
      red_add4(q, ar[SLOT], ag[SLOT], ab[SLOT], R.Y[0] * v);
      red_add4(q + 1, R.Y[1] * v, R.Y[2] * v, R.Y[3] * v, R.Y[4] * v);
      red_add4(q + 2, R.Y[5] * v, R.Y[6] * v, R.Y[7] * v, R.Y[8] * v);
    }
    ad[SLOT] = ar[SLOT] = ag[SLOT] = ab[SLOT] = aq[SLOT] = 0.0f;
  }
  // bring the gradient window to voxel (x,y,z): slots whose node changes are flushed first
  __device__ __forceinline__ void grad_window_move(int x, int y, int z)
  {
    const int v[3] = {x, y, z};
    int req[3][2];
    bool chg[3][2];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      req[a][0] = v[a] + (v[a] & 1);
      req[a][1] = v[a] + 1 - (v[a] & 1);
      chg[a][0] = ca[a][0] != req[a][0];
      chg[a][1] = ca[a][1] != req[a][1];
    }
    if (ca[0][0] >= 0) {   // not the very first sample
      if (chg[0][0] | chg[1][0] | chg[2][0]) flush_slot<0>();
      if (chg[0][1] | chg[1][0] | chg[2][0]) flush_slot<1>();
      if (chg[0][0] | chg[1][1] | chg[2][0]) flush_slot<2>();
      if (chg[0][1] | chg[1][1] | chg[2][0]) flush_slot<3>();
      if (chg[0][0] | chg[1][0] | chg[2][1]) flush_slot<4>();
      if (chg[0][1] | chg[1][0] | chg[2][1]) flush_slot<5>();
      if (chg[0][0] | chg[1][1] | chg[2][1]) flush_slot<6>();
      if (chg[0][1] | chg[1][1] | chg[2][1]) flush_slot<7>();
    }
#pragma unroll
    for (int a = 0; a < 3; a++) { ca[a][0] = req[a][0]; ca[a][1] = req[a][1]; }
  }
  __device__ __forceinline__ void flush_all()
  {
    if (ca[0][0] < 0) return;
    flush_slot<0>(); flush_slot<1>(); flush_slot<2>(); flush_slot<3>(); flush_slot<4>(); flush_slot<5>(); flush_slot<6>(); flush_slot<7>();
  }
  __device__ __forceinline__ bool sample(int x, int y, int z)
  {
    SampleGeom g;
    sample_geom(S, R, x, y, z, tPrev, g);
    window_density(S, W, x, y, z);
    const float dens_raw = window_density_lerp(W, g, x, y, z);
    const bool last = (k + 1 == n_total);
    if (dens_raw < 0.0f) {  // relu closed: sigma = 0, tr = exp(-0) = 1, weight 0 -> no adjoint reaches the grid (:170-171)
      k++;
      return !last;
    }
    const float sigma = dens_raw;
    const float tr = expf_libm(-sigma * g.delta, etab);
    const float wgt = thr * (1.0f - tr);
    float w[8];
    window_weights(g, x, y, z, w);
    // the colour is needed even when wgt == 0: d colour / d tr = -T_k c_k does not vanish with (1 - tr)
    float c[3] = {1.0f, 1.0f, 1.0f}, acc[3] = {0.0f, 0.0f, 0.0f}, raw = 0.0f, s = 0.0f;
    if (!grey) {
      window_colour(S, W, R.Y, x, y, z);
      window_colour_eval(W, w, acc, raw, s, c);
    }
    if (wgt != 0.0f) {
#pragma unroll
      for (int ch = 0; ch < 3; ch++) cpre[ch] += (double)(wgt * c[ch]);
    }
    const float dot = c[0] * aC[0] + c[1] * aC[1] + c[2] * aC[2];
    float a_tr;
    if (last) a_tr = -thr * dot + aTn * thr;
    else {
      const double suffix = (ctot[0] - cpre[0]) * (double)aC[0] + (ctot[1] - cpre[1]) * (double)aC[1] + (ctot[2] - cpre[2]) * (double)aC[2];
      a_tr = -thr * dot + (float)((suffix + (double)aTn * (double)Tn) / (double)tr);
    }
    // d tr / d sigma = -delta * tr ;  cauchy: d/dsigma log(1 + 2 sigma^2) = 4 sigma / (1 + 2 sigma^2), weight 0.0005 (:267)
    const float a_sigma = -g.delta * tr * a_tr + 0.0005f * (4.0f * sigma / (1.0f + 2.0f * sigma * sigma));
    const float a_dens = a_sigma;                              // relu passes iff density >= 0 (:170)
    float a_rgb[3] = {0.0f, 0.0f, 0.0f}, a_raw = 0.0f;
    if (!grey && wgt != 0.0f) {
      float a_s = 0.0f;
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const float a_c = wgt * aC[ch];
        const float u = acc[ch] * s;
        if (0.0f < u && u <= 1.0f) { a_rgb[ch] = s * a_c; a_s = fmaf(acc[ch], a_c, a_s); }   // clamp passes iff 0 < u <= 1
      }
      a_raw = (0.0f < raw && raw <= 1.0f) ? a_s : 0.0f;
    }
    if (a_dens != 0.0f || a_raw != 0.0f || a_rgb[0] != 0.0f || a_rgb[1] != 0.0f || a_rgb[2] != 0.0f) {
      grad_window_move(x, y, z);
#pragma unroll
      for (int cn = 0; cn < 8; cn++) {
        const float ww = w[cn];
        ad[cn] = fmaf(ww, a_dens, ad[cn]);
        ar[cn] = fmaf(ww, a_rgb[0], ar[cn]); ag[cn] = fmaf(ww, a_rgb[1], ag[cn]); ab[cn] = fmaf(ww, a_rgb[2], ab[cn]);
        aq[cn] = fmaf(ww, a_raw, aq[cn]);
      }
    }
    thr *= tr;
    k++;
    return !last;
  }
};

// hit enumeration for the visitors above: the fp32-filtered exact walker of lrf_wray.cuh (same hit list as march(), a fraction of
// its fp64 work) for every ray it can take, the generic march for the rest (a zero / non-finite reciprocal direction component)
template <class Visitor> __device__ __forceinline__ void walk(const SceneDev &S, const float *o, const float *d, Visitor &vis)
{
  if (!fast_path_ok(d)) {
    march(S, o, d, vis);
    return;
  }
  RayWalker<true> W(S);
  if (!W.start(o, d)) return;
  bool more = W.next();
  while (more) {
    if (!vis.sample(W.cell[0], W.cell[1], W.cell[2])) return;
    more = W.next();
  }
}

// idx_list / idx_count (packed-ray form only): when given, thread i handles ray idx_list[i] for i < *idx_count — the
// fallback queue of the warp-per-ray kernel (rays with a zero / non-finite reciprocal direction component)
template <bool IMAGE>
__global__ void __launch_bounds__(kBlock) grad_kernel(const __grid_constant__ SceneDev S, uint32_t w, uint32_t h, size_t n_rays, const float *__restrict__ ro,
                                                      const float *__restrict__ rd, const float *__restrict__ ref_rgba, int grey,
                                                      const uint32_t *__restrict__ idx_list, const uint32_t *__restrict__ idx_count)
{
  __shared__ uint64_t etab[32];
  load_exp_table(etab);
  size_t ray;
  float o[3], d[3];
  if (IMAGE) {
    uint32_t px, py;
    if (idx_list) {   // the fallback queue of grad_image_fast_kernel: pixel indices
      const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
      if (i >= *idx_count) return;
      px = idx_list[i] % w; py = idx_list[i] / w;
    } else if (!tile_pixel(w, h, px, py)) return;
    ray = (size_t)py * w + px;
    gen_ray(S.view_inv, S.proj_inv, px, py, w, h, o, d);
  } else {
    ray = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (idx_list) {
      if (ray >= *idx_count) return;
      ray = idx_list[ray];
    }
    if (ray >= n_rays) return;
#pragma unroll
    for (int a = 0; a < 3; a++) { o[a] = ro[ray * 3 + a]; d[a] = rd[ray * 3 + a]; }
  }
  FwdAVisitor fa(S, etab, grey != 0);
  ray_ctx_init(fa.R, o, d, grey != 0);
  walk(S, o, d, fa);
  if (fa.n_shaded == 0) return;  // ref :688  empty hit list: no loss term
  const float4 rp = __ldg(reinterpret_cast<const float4 *>(ref_rgba) + ray);
  float ref[3] = {rp.x, rp.y, rp.z};
  if (grey) {  // ref :696-704
    const float gsc = 0.11f * rp.z + 0.59f * rp.y + 0.3f * rp.x;
    ref[0] = ref[1] = ref[2] = gsc;
  }
  BwdVisitor bw(S, etab, grey != 0);
  bw.R = fa.R;
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    const float craw = fa.col[ch];
    const float cc = lm_clamp(craw, 0.0f, 1.0f);             // :193
    const float diff = cc - ref[ch];                         // :266
    const float sg = diff < 0.0f ? -1.0f : 1.0f;
    bw.aC[ch] = (0.0f < craw && craw <= 1.0f) ? sg : 0.0f;
    bw.ctot[ch] = fa.csum[ch];
  }
  // d/dT of 0.0001f * (log(T + 1e-3) + log(1 - T + 1e-3))   (:267, double arithmetic there)
  bw.Tn = fa.thr;
  bw.aTn = (float)((double)0.0001f * (1.0 / ((double)fa.thr + 1e-3) - 1.0 / ((double)(1.0f - fa.thr) + 1e-3)));
  bw.n_total = fa.n_shaded;
  walk(S, o, d, bw);
  bw.flush_all();
}

// Whole-image backward on the forward's machinery (lrf_fast.cuh MODE 2 / MODE 3): thread per ray, pass A = the forward with fp64
// colour sums, pass B = the adjoint of every sample through the same incremental window, gradients accumulated per window slot
// in shared memory and flushed with red.global.add when the slot's node changes.  Rays the streamlined march cannot take go
// to a queue that grad_kernel<true> (generic march) works off afterwards.
template <bool SPARSE, bool GREY>
__global__ void __launch_bounds__(kBlock, 3 * 128 / kBlock) grad_image_fast_kernel(const __grid_constant__ SceneDev S, uint32_t w, uint32_t h,
                                                                    const float *__restrict__ ref_rgba, uint32_t *__restrict__ fb_count,
                                                                    uint32_t *__restrict__ fb_list)
{
  __shared__ uint64_t etab[32];
  __shared__ float ysm[9 * kFastBlock];
  __shared__ float4 csm[8 * kFastBlock];
  __shared__ float4 gsm[8 * kFastBlock];
  __shared__ float gdm[8 * kFastBlock];
  __shared__ uint32_t nsm[8 * kFastBlock];
  load_exp_table(etab);
  uint32_t px, py;
  if (!tile_pixel(w, h, px, py)) return;
  float o[3], d[3];
  gen_ray(S.view_inv, S.proj_inv, px, py, w, h, o, d);
  if (!fast_path_ok(d)) {
    fb_list[atomicAdd(fb_count, 1u)] = py * w + px;
    return;
  }
  double ctot[3];
  float craw[3], Tn;
  uint32_t n_total;
  {
    FastMarch<2, SPARSE, GREY> fa(S, etab, ysm + threadIdx.x, csm + threadIdx.x);
    fa.run(o, d);
    if (fa.n_shaded == 0) return;                              // ref :688  empty hit list: no loss term
    n_total = fa.n_shaded;
    Tn = fa.thr;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) { ctot[ch] = fa.csum[ch]; craw[ch] = fa.col[ch]; }
  }
  const float4 rp = __ldg(reinterpret_cast<const float4 *>(ref_rgba) + (size_t)py * w + px);
  float ref[3] = {rp.x, rp.y, rp.z};
  if (GREY) {  // ref :696-704
    const float gsc = 0.11f * rp.z + 0.59f * rp.y + 0.3f * rp.x;
    ref[0] = ref[1] = ref[2] = gsc;
  }
  FastMarch<3, SPARSE, GREY> fb(S, etab, ysm + threadIdx.x, csm + threadIdx.x);
  fb.gsm = gsm + threadIdx.x; fb.gdm = gdm + threadIdx.x; fb.nsm = nsm + threadIdx.x;
  fb.grad_window_reset();
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    const float cc = lm_clamp(craw[ch], 0.0f, 1.0f);           // :193
    const float diff = cc - ref[ch];                           // :266
    const float sg = diff < 0.0f ? -1.0f : 1.0f;
    fb.aC[ch] = (0.0f < craw[ch] && craw[ch] <= 1.0f) ? sg : 0.0f;
    fb.ctot[ch] = ctot[ch];
  }
  fb.Tn = Tn;
  // d/dT of 0.0001f * (log(T + 1e-3) + log(1 - T + 1e-3))   (:267, double arithmetic there)
  fb.aTn = (float)((double)0.0001f * (1.0 / ((double)Tn + 1e-3) - 1.0 / ((double)(1.0f - Tn) + 1e-3)));
  fb.n_total = n_total;
  fb.run(o, d);
  fb.grad_flush_all();
}

// packed ray batch, one WARP per ray (lrf_wray.cuh)
// Block shape of the warp-per-ray kernel.  All rays of a batch take about the same time, so the batch runs in whole "waves":
// 4096 rays on 148 SMs x W resident warps.  W = 12 (168 registers, no spills) is 2.3 waves = three rounds with the last one
// 30 % full; the register file is split over 4 schedulers, so the next step is W = 16 at 128 registers (158 bytes of spills):
// 1.73 waves, two rounds.  Measured 0.452 ms (64 threads x 7 or 8 blocks) against 0.463 ms (128 x 3): kept, a 2.4 % gain.
#ifndef LRF_WRAY_BLOCK
#define LRF_WRAY_BLOCK 64
#endif
#ifndef LRF_WRAY_MINB
#define LRF_WRAY_MINB 7
#endif
constexpr int kWrayBlock = LRF_WRAY_BLOCK;
template <bool SPARSE, bool GREY>
__global__ void __launch_bounds__(kWrayBlock, LRF_WRAY_MINB) grad_rays_warp_kernel(const __grid_constant__ SceneDev S, size_t n_rays, const float *__restrict__ ro,
                                                                const float *__restrict__ rd, const float *__restrict__ ref_rgba,
                                                                uint32_t *__restrict__ fb_count, uint32_t *__restrict__ fb_list,
                                                                HitRec *__restrict__ hitbuf, uint32_t hitcap, int parallel_gather)
{
  __shared__ uint64_t etab[32];
  __shared__ uint32_t wslots[kWrayBlock / 32][36];             // per warp: the rank -> crossing-count table of WarpRay::gather_parallel
  load_exp_table(etab);
  const size_t ray = ((size_t)blockIdx.x * kWrayBlock + threadIdx.x) >> 5;
  if (ray >= n_rays) return;
  float o[3], d[3];
#pragma unroll
  for (int a = 0; a < 3; a++) { o[a] = __ldg(ro + ray * 3 + a); d[a] = __ldg(rd + ray * 3 + a); }
  if (!fast_path_ok(d)) {
    if ((threadIdx.x & 31) == 0) fb_list[atomicAdd(fb_count, 1u)] = (uint32_t)ray;
    return;
  }
  const float4 rp = __ldg(reinterpret_cast<const float4 *>(ref_rgba) + ray);
  float ref[3] = {rp.x, rp.y, rp.z};
  if (GREY) {  // ref :696-704
    const float gsc = 0.11f * rp.z + 0.59f * rp.y + 0.3f * rp.x;
    ref[0] = ref[1] = ref[2] = gsc;
  }
  warp_ray_backward<SPARSE, GREY>(S, etab, o, d, ref, hitbuf ? hitbuf + ray * hitcap : nullptr, hitcap, parallel_gather ? wslots[threadIdx.x >> 5] : nullptr);
}

// device-side ray generation for (view, pixel) picks — ref :635-639
__global__ void gen_rays_kernel(size_t n, const float *__restrict__ view_inv, const uint32_t *__restrict__ picks, uint32_t w, uint32_t h,
                                float *__restrict__ o, float *__restrict__ d, const __grid_constant__ SceneDev S)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t view = picks[2 * i], pix = picks[2 * i + 1];
  float vi[16];
#pragma unroll
  for (int j = 0; j < 16; j++) vi[j] = __ldg(view_inv + (size_t)view * 16 + j);
  float oo[3], dd[3];
  gen_ray(vi, S.proj_inv, pix % w, pix / w, w, h, oo, dd);
#pragma unroll
  for (int a = 0; a < 3; a++) { o[i * 3 + a] = oo[a]; d[i * 3 + a] = dd[a]; }
}

// ------------------------------------------------------------------------------------------------
// Synthetic grids of SURVEY 8(d), generated in place (a 512^3 grid is 7 GB: it never exists on the host).  Same values,
// bit for bit, as the host generator tests/libs.py::lcg_grid: 64-bit LCG s <- s*A + C from the fixed seed, 13 draws per node
// in node order, u = (s >> 40) / 2^24.  Thread i jumps the LCG ahead by 13*i steps (square-and-multiply on the affine map).
//   mode 0: density = D*(u - 0.25), RGB = u, sh0 = 1 + 2u, sh1..8 = u - 0.5          ("opaque" D = 20 / "thin" D = 2)
//   mode 1: mode 0 with the density multiplied by the Gaussian blob exp(-|p - .5|^2 / (2 * 0.2^2))
//   mode 2: |mode-0 density| + 0.5 inside the sphere of radius 0.288 about the centre (volume fraction 0.100), minus that
//           outside — BASELINE config 4's sparse scene: the box criterion (centre density > 0) keeps ~10 % of the voxels
__global__ void fill_synthetic_kernel(int gs, float D, int mode, float *__restrict__ dens, float4 *__restrict__ col)
{
  const size_t N = (size_t)gs * gs * gs, i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint64_t A = 6364136223846793005ULL, Cc = 1442695040888963407ULL;
  // affine map after k steps: s -> a*s + c, built by doubling
  uint64_t a = 1, c = 0, pa = A, pc = Cc;
  for (uint64_t k = 13ULL * i; k; k >>= 1) {
    if (k & 1) { a = pa * a; c = pa * c + pc; }
    pc = pa * pc + pc;
    pa = pa * pa;
  }
  uint64_t st = a * 0x9E3779B97F4A7C15ULL + c;
  float u[13];
#pragma unroll
  for (int j = 0; j < 13; j++) {
    st = st * A + Cc;
    u[j] = (float)(st >> 40) / 16777216.0f;
  }
  float d0 = D * (u[0] - 0.25f);
  if (mode != 0) {
    const int x = (int)(i % gs), y = (int)((i / gs) % gs), z = (int)(i / ((size_t)gs * gs));
    const float px = ((float)x + 0.5f) / (float)gs - 0.5f, py = ((float)y + 0.5f) / (float)gs - 0.5f, pz = ((float)z + 0.5f) / (float)gs - 0.5f;
    const float r2 = px * px + py * py + pz * pz;
    if (mode == 1) d0 *= (float)exp((double)(-(r2) / (float)(2.0 * 0.2 * 0.2)));   // numpy float32 exp == correctly rounded here: tests pin it
    else d0 = (r2 < 0.288f * 0.288f) ? (fabsf(d0) + 0.5f) : -(fabsf(d0) + 0.5f);
  }
  dens[i] = d0;
  col[i * 3] = make_float4(u[1], u[2], u[3], 1.0f + 2.0f * u[4]);
  col[i * 3 + 1] = make_float4(u[5] - 0.5f, u[6] - 0.5f, u[7] - 0.5f, u[8] - 0.5f);
  col[i * 3 + 2] = make_float4(u[9] - 0.5f, u[10] - 0.5f, u[11] - 0.5f, u[12] - 0.5f);
}
// the InitGrid box list (gs-1)^3 in id order (ref example_tracer.h:69-85), materialised on the device
__global__ void dense_boxes_kernel(int n, int32_t *__restrict__ xyz)
{
  const size_t total = (size_t)n * n * n, i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  xyz[3 * i] = (int32_t)(i % n);
  xyz[3 * i + 1] = (int32_t)((i / n) % n);
  xyz[3 * i + 2] = (int32_t)(i / ((size_t)n * n));
}

// ------------------------------------------------------------------------------------------------
// Callers either side of the path.
// Target ingest — what stbi_loadf does to a decoded 8-bit RGBA image (ref main.cpp:72,96; external/stb_image.h:1856-1876
// stbi__ldr_to_hdr: colour = (float)(pow(u8/255.0f, 2.2f) * 1.0f), alpha = u8/255.0f; :1241-1258 vertical flip).  The 256
// possible colour values are tabulated on the HOST with the same libm call, so the device result is that value by
// construction; the alpha division is a correctly rounded fp32 divide on both sides.
__global__ void ingest_rgba8_kernel(const uchar4 *__restrict__ src, float4 *__restrict__ dst, uint32_t w, uint32_t h, int flip,
                                    const float *__restrict__ lut)
{
  __shared__ float slut[256];
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) slut[i] = lut[i];
  __syncthreads();
  const size_t n = (size_t)w * h;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i % w);
    const uchar4 p = src[(size_t)(flip ? h - 1 - y : y) * w + x];
    dst[i] = make_float4(slut[p.x], slut[p.y], slut[p.z], (float)p.w / 255.0f);
  }
}
// reference pixels of a batch of (view, pixel) picks out of the resident targets (config 3's random-ray batches)
__global__ void gather_targets_kernel(size_t n, const uint32_t *__restrict__ picks, const float4 *__restrict__ targets, size_t view_stride,
                                      float4 *__restrict__ out)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = __ldg(targets + (size_t)picks[2 * i] * view_stride + picks[2 * i + 1]);
}
// BMP epilogue — LiteImage::SaveBMP's pixel conversion (ref external/LiteMath/Image2d.cpp:519-533): the 24-bit record is
// {(p >> 16) & 255, (p >> 8) & 255, p & 255} in file order (a BMP record is blue, green, red: the packed frame's B byte
// goes first).  No row padding: the reference writes width*height*3 bytes whatever the width.
__global__ void pack_bmp_kernel(const uint32_t *__restrict__ pix, uint8_t *__restrict__ out, size_t n)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t p = pix[i];
    out[3 * i] = (uint8_t)((p & 0x00FF0000u) >> 16);
    out[3 * i + 1] = (uint8_t)((p & 0x0000FF00u) >> 8);
    out[3 * i + 2] = (uint8_t)(p & 0x000000FFu);
  }
}

// ------------------------------------------------------------------------------------------------
// TV (density only; ref TVRegularisation :196-239, the SH loop :241-245 is dead — SURVEY §2 defects)
// stage 1: V_raw(a,b,c) = trilinear(0.5) of density over {a,min(a+1,gs-1)} x ... (getVoxel :203-230), pre-relu
__global__ void tv_voxel_kernel(int gs, const float *__restrict__ dens, float *__restrict__ vraw)
{
  const size_t N = (size_t)gs * gs * gs, i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int a = (int)(i % gs), b = (int)((i / gs) % gs), c = (int)(i / ((size_t)gs * gs));
  const int a1 = min(a + 1, gs - 1), b1 = min(b + 1, gs - 1), c1 = min(c + 1, gs - 1);
  auto at = [&](int x, int y, int z) { return __ldg(dens + (size_t)x + (size_t)y * gs + (size_t)z * gs * gs); };
  const float xy00 = lm_lerp(at(a, b, c), at(a1, b, c), 0.5f), xy10 = lm_lerp(at(a, b1, c), at(a1, b1, c), 0.5f);
  const float xy01 = lm_lerp(at(a, b, c1), at(a1, b, c1), 0.5f), xy11 = lm_lerp(at(a, b1, c1), at(a1, b1, c1), 0.5f);
  vraw[i] = lm_lerp(lm_lerp(xy00, xy10, 0.5f), lm_lerp(xy01, xy11, 0.5f), 0.5f);
}
// stage 2: G(p) = d(sum tv)/dV(p) gathered from the <= 4 tv cells that reference V(p), relu-gated
__global__ void tv_gather_kernel(int gs, const float *__restrict__ vraw, float *__restrict__ gv)
{
  const size_t N = (size_t)gs * gs * gs, i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int a = (int)(i % gs), b = (int)((i / gs) % gs), c = (int)(i / ((size_t)gs * gs));
  auto V = [&](int x, int y, int z) { const float v = __ldg(vraw + (size_t)x + (size_t)y * gs + (size_t)z * gs * gs); return v < 0.0f ? 0.0f : v; };
  const float lam = 0.01f;
  // contribution of tv cell (x,y,z) (requires x,y,z < gs-1): which = 0 centre, 1/2/3 = it references us as its +x/+y/+z neighbour
  auto cell = [&](int x, int y, int z, int which) -> float {
    const float v = V(x, y, z);
    const float dx = v - V(x + 1, y, z), dy = v - V(x, y + 1, z), dz = v - V(x, y, z + 1);
    const float sum = dx * dx + dy * dy + dz * dz;
    if (!(sum > 0.0f)) return 0.0f;  // sqrt'(0) := 0
    const float r = sqrtf(sum);
    const float num = which == 0 ? (dx + dy + dz) : which == 1 ? -dx : which == 2 ? -dy : -dz;
    return lam * num / r;
  };
  float g = 0.0f;
  const int m = gs - 1;
  if (a < m && b < m && c < m) g += cell(a, b, c, 0);
  if (a >= 1 && b < m && c < m) g += cell(a - 1, b, c, 1);
  if (b >= 1 && a < m && c < m) g += cell(a, b - 1, c, 2);
  if (c >= 1 && a < m && b < m) g += cell(a, b, c - 1, 3);
  gv[i] = (__ldg(vraw + i) < 0.0f) ? 0.0f : g;  // relu gate of getVoxel (:226-227)
}
// stage 3: density-node gradient += sum over the V cells having the node as a corner of (0.5 per axis, doubled where far == near)
__global__ void tv_scatter_kernel(int gs, const float *__restrict__ gv, float *__restrict__ gdens)
{
  const size_t N = (size_t)gs * gs * gs, i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int q[3] = {(int)(i % gs), (int)((i / gs) % gs), (int)(i / ((size_t)gs * gs))};
  float acc = 0.0f;
  for (int dz = -1; dz <= 0; dz++)
    for (int dy = -1; dy <= 0; dy++)
      for (int dx = -1; dx <= 0; dx++) {
        const int p[3] = {q[0] + dx, q[1] + dy, q[2] + dz};
        if (p[0] < 0 || p[1] < 0 || p[2] < 0) continue;
        float wgt = 1.0f;
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
          const int mult = (q[ax] == p[ax] ? 1 : 0) + (q[ax] == min(p[ax] + 1, gs - 1) ? 1 : 0);
          wgt *= 0.5f * (float)mult;
        }
        acc += wgt * __ldg(gv + (size_t)p[0] + (size_t)p[1] * gs + (size_t)p[2] * gs * gs);
      }
  gdens[i] += acc;
}

// Fused TV gradient: the three stages above in ONE kernel over 8^3 output tiles staged through shared memory
// (12^3 densities -> 11^3 V -> 10^3 tv cells -> 9^3 G -> 8^3 node gradients).  The stage-by-stage version moves the
// V and G planes through L2 twice and evaluates every tv cell four times from global memory; here the only global
// traffic is the density halo in and the density-gradient tile in/out, and every tv cell (the sqrt) is evaluated once
// per tile.  Values agree with the staged kernels to rounding (lam * d / r is formed as d * (lam * rsqrt)); the TV
// gradient is compared with the oracle to 1e-5 relative, not bit-for-bit (the reference's own sum order is a race).
constexpr int kTvT = 8;
// (x, y, z) of a linear index into an A^3 cube, advanced by the block's stride of 256 without divisions: ncu showed the first
// version of tv_fused_kernel issue-bound (87 % issue slots busy) with 55 % of its instructions integer index arithmetic
// (i % A, i / A % A, i / A^2 per element); the decomposition is now done once per thread and carried incrementally.
template <int A> struct CubeIdx
{
  int x, y, z;
  __device__ __forceinline__ explicit CubeIdx(int i) : x(i % A), y((i / A) % A), z(i / (A * A)) {}
  __device__ __forceinline__ void step()
  {
    constexpr int sx = 256 % A, sy = (256 / A) % A, sz = 256 / (A * A);
    x += sx;
    int carry = 0;
    if (x >= A) { x -= A; carry = 1; }
    y += sy + carry;
    carry = 0;
    if (y >= A) { y -= A; carry = 1; }
    z += sz + carry;
  }
};
__global__ void __launch_bounds__(256) tv_fused_kernel(int gs, const float *__restrict__ dens, float *__restrict__ gdens, int bz0, int atomic,
                                                       int n_tiles)
{
  __shared__ float sD[12 * 12 * 12], sV[11 * 11 * 11], sG[9 * 9 * 9];
  __shared__ float sEx[10 * 10 * 10], sEy[10 * 10 * 10], sEz[10 * 10 * 10];
  const int tiles = gs / kTvT;
  const int m = gs - 1;
  // a block walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ...: with gridDim.x == n_tiles one tile per block; a small
  // grid (2 blocks per SM) is used when the kernel shares the GPU with the HBM-bound Adam pass of the colour planes
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  const int bx = tile % tiles, by = (tile / tiles) % tiles, bz = bz0 + tile / (tiles * tiles);   // bz0: first tile layer of a z-slab
  const int q0x = bx * kTvT, q0y = by * kTvT, q0z = bz * kTvT;
  CubeIdx<12> c12(threadIdx.x);
  for (int i = threadIdx.x; i < 12 * 12 * 12; i += 256, c12.step()) {
    const int lx = c12.x, ly = c12.y, lz = c12.z;
    const int x = min(max(q0x - 2 + lx, 0), m), y = min(max(q0y - 2 + ly, 0), m), z = min(max(q0z - 2 + lz, 0), m);
    sD[i] = __ldg(dens + (size_t)x + (size_t)y * gs + (size_t)z * gs * gs);
  }
  __syncthreads();
  // V_raw at local l <-> coordinate q0 - 2 + l: trilinear(0.5) over {a, min(a+1, gs-1)}^3 (getVoxel :203-230)
  CubeIdx<11> c11(threadIdx.x);
  for (int i = threadIdx.x; i < 11 * 11 * 11; i += 256, c11.step()) {
    const int lx = c11.x, ly = c11.y, lz = c11.z;
    const float *p = sD + lx + ly * 12 + lz * 144;
    const float xy00 = lm_lerp(p[0], p[1], 0.5f), xy10 = lm_lerp(p[12], p[13], 0.5f);
    const float xy01 = lm_lerp(p[144], p[145], 0.5f), xy11 = lm_lerp(p[156], p[157], 0.5f);
    sV[i] = lm_lerp(lm_lerp(xy00, xy10, 0.5f), lm_lerp(xy01, xy11, 0.5f), 0.5f);
  }
  __syncthreads();
  // tv cell at local l <-> coordinate q0 - 2 + l (needs V at l and l + 1): E = lam * (dx, dy, dz) / sqrt(dx^2 + dy^2 + dz^2),
  // zero where the cell does not exist (a coordinate outside [0, gs-2]) or the root's argument is 0 (sqrt'(0) := 0)
  CubeIdx<10> c10(threadIdx.x);
  for (int i = threadIdx.x; i < 10 * 10 * 10; i += 256, c10.step()) {
    const int lx = c10.x, ly = c10.y, lz = c10.z;
    const int a = q0x - 2 + lx, b = q0y - 2 + ly, c = q0z - 2 + lz;
    float ex = 0.0f, ey = 0.0f, ez = 0.0f;
    if (a >= 0 && b >= 0 && c >= 0 && a < m && b < m && c < m) {
      const float *p = sV + lx + ly * 11 + lz * 121;
      const float v = fmaxf(p[0], 0.0f);
      const float dx = v - fmaxf(p[1], 0.0f), dy = v - fmaxf(p[11], 0.0f), dz = v - fmaxf(p[121], 0.0f);
      const float sum = dx * dx + dy * dy + dz * dz;
      if (sum > 0.0f) {
        const float k = 0.01f * rsqrtf(sum);
        ex = dx * k; ey = dy * k; ez = dz * k;
      }
    }
    sEx[i] = ex; sEy[i] = ey; sEz[i] = ez;
  }
  __syncthreads();
  // G(p) = d(sum tv)/dV(p), p = q0 - 1 + l: + (Ex+Ey+Ez)(p) - Ex(p - ex) - Ey(p - ey) - Ez(p - ez), relu-gated (:226-227)
  CubeIdx<9> c9(threadIdx.x);
  for (int i = threadIdx.x; i < 9 * 9 * 9; i += 256, c9.step()) {
    const int lx = c9.x, ly = c9.y, lz = c9.z;
    const int e = (lx + 1) + (ly + 1) * 10 + (lz + 1) * 100;   // tv-cell index of p
    float g = sEx[e] + sEy[e] + sEz[e] - sEx[e - 1] - sEy[e - 10] - sEz[e - 100];
    if (sV[(lx + 1) + (ly + 1) * 11 + (lz + 1) * 121] < 0.0f) g = 0.0f;
    if (q0x - 1 + lx < 0 || q0y - 1 + ly < 0 || q0z - 1 + lz < 0) g = 0.0f;
    sG[i] = g;
  }
  __syncthreads();
  // node gradient: each node is a corner of up to 8 V cells, weight 0.5 per axis, doubled where far == near (last node row)
  for (int i = threadIdx.x; i < kTvT * kTvT * kTvT; i += 256) {
    const int lx = i % kTvT, ly = (i / kTvT) % kTvT, lz = i / (kTvT * kTvT);
    const int qx = q0x + lx, qy = q0y + ly, qz = q0z + lz;
    // along an axis: contribution of V cell q (weight 0.5, or 1.0 at the last node where min(q+1, gs-1) == q) and of
    // V cell q - 1 (weight 0.5, absent at q == 0)
    const float wx0 = qx == m ? 1.0f : 0.5f, wx1 = qx > 0 ? 0.5f : 0.0f;
    const float wy0 = qy == m ? 1.0f : 0.5f, wy1 = qy > 0 ? 0.5f : 0.0f;
    const float wz0 = qz == m ? 1.0f : 0.5f, wz1 = qz > 0 ? 0.5f : 0.0f;
    const float *p = sG + (lx + 1) + (ly + 1) * 9 + (lz + 1) * 81;
    float acc = 0.0f;
    acc += wz1 * (wy1 * (wx1 * p[-81 - 9 - 1] + wx0 * p[-81 - 9]) + wy0 * (wx1 * p[-81 - 1] + wx0 * p[-81]));
    acc += wz0 * (wy1 * (wx1 * p[-9 - 1] + wx0 * p[-9]) + wy0 * (wx1 * p[-1] + wx0 * p[0]));
    float *dst = gdens + (size_t)qx + (size_t)qy * gs + (size_t)qz * gs * gs;
    if (atomic) red_add_f32(dst, acc);   // running beside the ray backward (which scatters with red.add into the same plane)
    else *dst += acc;
  }
  __syncthreads();   // the shared arrays are reused by the next tile
  }
}

// ------------------------------------------------------------------------------------------------
// Adam, ref optimizerStep example_tracer.h:152-162: fp32 moments, fp64 update expression
//     V = b1 V + (1-b1) g;  Vh = V / (1 - b1^t);  S = b2 S + (1-b2) g g;  Sh = S / (1 - b2^t);  p -= lr Vh / (sqrt(Sh) + eps)
// with Vh, Sh and the whole update in double, the result narrowed to float.  Evaluated literally that is 3 IEEE fp64
// divisions and an fp64 sqrt per element (~120 fp64-pipe instructions): the first version of this kernel ran at 50 %
// of the HBM roofline, bound by the fp64 pipe.  The result is a FLOAT, though, so the double value only has to be known
// well enough to round it: update() evaluates the expression with reciprocal / rsqrt seeds + one Newton step each
// (relative error < 2^-40, ~30 instructions) and accepts the float only if the whole error interval rounds to the same
// float; otherwise (about one element in 10^5) the literal IEEE sequence decides.  Bit-exact either way.
struct AdamConst
{
  float lr, beta1, beta2, eps, omb1f, omb2f;
  double om_b1, om_b2, inv_om_b1, inv_om_b2;
};
__device__ __forceinline__ float adam_param_exact(float p, float v, float s, const AdamConst &C)
{
  const double Vh = (double)v / C.om_b1;
  const double Sh = (double)s / C.om_b2;
  return (float)((double)p - (double)C.lr * Vh / (sqrt(Sh) + (double)C.eps));
}
__device__ __forceinline__ void adam_update(float &p, float g, float &v, float &s, const AdamConst &C)
{
  v = C.beta1 * v + C.omb1f * g;
  s = C.beta2 * s + C.omb2f * g * g;
  if (v == 0.0f && s == 0.0f) return;   // never touched by a gradient: lr * 0 / (0 + eps) = 0, p - 0 = p exactly (most of the grid early on)
  // fast evaluation: every approximate step carries a relative error far below 2^-40
  const double Sh = (double)s * C.inv_om_b2;
  const float shf = (float)Sh;
  bool ok = (shf > 1e-30f) && (shf < 1e30f);
  const double y0 = (double)rsqrtf(shf);                       // 2^-22
  const double y1 = y0 * fma(-0.5 * Sh, y0 * y0, 1.5);         // one Newton step on 1/sqrt: ~2^-43
  const double D = fma(Sh, y1, (double)C.eps);                 // sqrt(Sh) + eps
  const float df = (float)D;
  ok = ok && (df > 1e-30f) && (df < 1e30f);
  const double r0 = (double)__fdividef(1.0f, df);              // rcp.approx: 2^-23
  const double r1 = r0 * fma(-D, r0, 2.0);                     // one Newton step on 1/D: ~2^-46
  const double x = ((double)C.lr * ((double)v * C.inv_om_b1)) * r1;
  const double r = (double)p - x;
  const double e = fabs(x) * 0x1p-38 + fabs(r) * 0x1p-50;      // generous bound on |r - r_reference|
  const float lo = (float)(r - e), hi = (float)(r + e);
  if (ok && lo == hi) p = lo;
  else p = adam_param_exact(p, v, s, C);
}

__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ V, float *__restrict__ Sq,
                                                   size_t count, const AdamConst C, int zero_grad)
{
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= count) return;
  if (i4 + 4 <= count) {
    float4 pv = *reinterpret_cast<float4 *>(p + i4), gv = *reinterpret_cast<float4 *>(g + i4);
    float4 vv = *reinterpret_cast<float4 *>(V + i4), sv = *reinterpret_cast<float4 *>(Sq + i4);
    adam_update(pv.x, gv.x, vv.x, sv.x, C); adam_update(pv.y, gv.y, vv.y, sv.y, C);
    adam_update(pv.z, gv.z, vv.z, sv.z, C); adam_update(pv.w, gv.w, vv.w, sv.w, C);
    *reinterpret_cast<float4 *>(p + i4) = pv;
    *reinterpret_cast<float4 *>(V + i4) = vv;
    *reinterpret_cast<float4 *>(Sq + i4) = sv;
    if (zero_grad) *reinterpret_cast<float4 *>(g + i4) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (size_t i = i4; i < count; i++) {
      float pp = p[i], vv = V[i], ss = Sq[i];
      adam_update(pp, g[i], vv, ss, C);
      p[i] = pp; V[i] = vv; Sq[i] = ss;
      if (zero_grad) g[i] = 0.0f;
    }
  }
}

// Multi-GPU exchange step fused with the per-cell pass, over NVLink peer memory: every rank owns a slab of
// the element range.  For its slab it reads the gradient of ALL ranks (its own from HBM, the others' through peer
// loads), sums them in rank order — the full-batch gradient, exactly what a sum all-reduce would leave there —, applies
// Adam with its local moments, and stores the new parameters into EVERY replica (peer stores).  One pass replaces
// reduce-scatter + Adam + all-gather: per rank (n-1)/n of the gradient comes in and (n-1)/n of the parameters go out over
// NVLink, the same volume as the collectives, but nothing is staged and the Adam arithmetic is divided by n.
// Ordering across GPUs is the caller's: a barrier before (all gradients complete) and after (all replicas written).
struct PeerSet
{
  int world, rank;
  float *grad[8];
  float *param[8];
};
__device__ __forceinline__ float4 ld_peer(const float *p)
{
  // peer memory is another GPU's HBM.  .cg = no L1 allocation (nothing is re-read), requests still coalesce into full lines
  // (.cv — "fetch again every time" — splits them and costs NVLink packet efficiency)
  return __ldcg(reinterpret_cast<const float4 *>(p));
}
__global__ void __launch_bounds__(256) adam_peer_kernel(const PeerSet P, float *__restrict__ V, float *__restrict__ Sq, size_t off, size_t count,
                                                        const AdamConst C)
{
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= count) return;
  const size_t at = off + i4;
  if (i4 + 4 <= count) {
    // all ranks' loads are issued before the first one is consumed (a peer load is a microsecond away); summed in rank order
    float4 gr[8];
#pragma unroll
    for (int r = 0; r < 8; r++)
      if (r < P.world) gr[r] = ld_peer(P.grad[r] + at);
    float4 pv = *reinterpret_cast<const float4 *>(P.param[P.rank] + at);
    float4 vv = *reinterpret_cast<float4 *>(V + at), sv = *reinterpret_cast<float4 *>(Sq + at);
    float4 gv = gr[0];
#pragma unroll
    for (int r = 1; r < 8; r++)
      if (r < P.world) { gv.x += gr[r].x; gv.y += gr[r].y; gv.z += gr[r].z; gv.w += gr[r].w; }
    adam_update(pv.x, gv.x, vv.x, sv.x, C); adam_update(pv.y, gv.y, vv.y, sv.y, C);
    adam_update(pv.z, gv.z, vv.z, sv.z, C); adam_update(pv.w, gv.w, vv.w, sv.w, C);
    *reinterpret_cast<float4 *>(V + at) = vv;
    *reinterpret_cast<float4 *>(Sq + at) = sv;
#pragma unroll
    for (int r = 0; r < 8; r++)
      if (r < P.world) *reinterpret_cast<float4 *>(P.param[r] + at) = pv;
  } else {
    for (size_t i = at; i < off + count; i++) {
      float g = P.grad[0][i];
      for (int r = 1; r < P.world; r++) g += P.grad[r][i];
      float pp = P.param[P.rank][i], vv = V[i], ss = Sq[i];
      adam_update(pp, g, vv, ss, C);
      V[i] = vv; Sq[i] = ss;
      for (int r = 0; r < P.world; r++) P.param[r][i] = pp;
    }
  }
}

// The same exchange through the NVSwitch (NVLink multicast objects, "NVLS"): the gradient planes of all ranks are bound to ONE
// multicast address range and so are the parameter planes.  multimem.ld_reduce asks the switch for the SUM of every replica's value
// at an address (one reduced 16-byte response comes back instead of world - 1 peer loads), multimem.st stores one value into every
// replica (one store goes out instead of world - 1 peer stores): per rank and direction the NVLink volume falls from
// 2 (n-1)/n to (n-1)/n + 1/n of a plane set.  The switch adds in an order of its own, so the sum may differ from the rank-order sum
// of adam_peer_kernel in the last bit; every element is reduced once (by its owner) and broadcast, so the replicas stay identical.
__device__ __forceinline__ float4 mc_ld_reduce(const float *mc)
{
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float *mc, const float4 &v)
{
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// count is a multiple of 4 (slabs of whole z planes of a grid whose size is a multiple of 8); UNROLL float4 per thread, all
// reductions requested before the first one is consumed (a switch round trip is microseconds)
template <int UNROLL>
__global__ void __launch_bounds__(256) adam_mc_kernel(const float *__restrict__ mc_grad, float *__restrict__ mc_param, const float *__restrict__ param,
                                                      float *__restrict__ V, float *__restrict__ Sq, size_t off, size_t count, const AdamConst C)
{
  // grid-stride over chunks of 256 * UNROLL float4: a persistent grid (launch_adam_mc caps it) keeps every SM's request queues full
  for (size_t chunk = blockIdx.x; chunk * (256 * UNROLL * 4) < count; chunk += gridDim.x) {
    const size_t base = (chunk * 256 * UNROLL + threadIdx.x) * 4;
    float4 g[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t i4 = base + (size_t)u * 256 * 4;
      if (i4 < count) g[u] = mc_ld_reduce(mc_grad + off + i4);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t i4 = base + (size_t)u * 256 * 4;
      if (i4 >= count) continue;
      const size_t at = off + i4;
      float4 pv = *reinterpret_cast<const float4 *>(param + at);
      float4 vv = *reinterpret_cast<float4 *>(V + at), sv = *reinterpret_cast<float4 *>(Sq + at);
      const float4 gv = g[u];
      adam_update(pv.x, gv.x, vv.x, sv.x, C); adam_update(pv.y, gv.y, vv.y, sv.y, C);
      adam_update(pv.z, gv.z, vv.z, sv.z, C); adam_update(pv.w, gv.w, vv.w, sv.w, C);
      *reinterpret_cast<float4 *>(V + at) = vv;
      *reinterpret_cast<float4 *>(Sq + at) = sv;
      mc_st(mc_param + at, pv);
    }
  }
}

// Inter-GPU barrier over peer memory for the exchange above: rank r stores the epoch into slot r of EVERY rank's flag array
// (release, system scope) and waits until every slot of its own array has reached the epoch (acquire, system scope).  Stream
// order makes the kernels queued before it complete first, so "all ranks passed the barrier" implies "all ranks finished what
// they had queued before it".  ~10 us on an NVSwitch box, against ~170 us for a 1-element NCCL all-reduce at 8 GPUs.
// A rank that never arrives (a crashed peer) trips the time-out instead of hanging the GPU: *timed_out is set and the wait ends.
struct PeerFlags
{
  int world, rank;
  uint32_t *flags[8];
};
__global__ void peer_barrier_kernel(const PeerFlags F, uint32_t epoch, uint32_t *timed_out)
{
  const int t = threadIdx.x;
  if (t >= F.world) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(F.flags[t] + F.rank), "r"(epoch) : "memory");
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(F.flags[F.rank] + t) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 2000000000ULL) {   // 2 s
      *timed_out = 1u;
      break;
    }
  }
}

__global__ void fill_kernel(float *__restrict__ p, size_t count, float v)
{
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 4 <= count) *reinterpret_cast<float4 *>(p + i4) = make_float4(v, v, v, v);
  else
    for (size_t i = i4; i < count; i++) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// layout conversion between the reference's AoS Cell[13] and the device planes
__global__ void aos_to_planes_kernel(const float *__restrict__ aos, float *__restrict__ planes, size_t N)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * kCell) return;
  const size_t node = t / kCell;
  const int j = (int)(t % kCell);
  const float v = aos[t];
  if (j == 0) planes[node] = v;
  else planes[N + node * 12 + (j - 1)] = v;
}
__global__ void planes_to_aos_kernel(const float *__restrict__ planes, float *__restrict__ aos, size_t N)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * kCell) return;
  const size_t node = t / kCell;
  const int j = (int)(t % kCell);
  aos[t] = (j == 0) ? planes[node] : planes[N + node * 12 + (j - 1)];
}

// ------------------------------------------------------------------------------------------------
// level transition
__device__ __forceinline__ float plane_get(const float *planes, size_t N, size_t node, int j)
{
  return j == 0 ? __ldg(planes + node) : __ldg(planes + N + node * 12 + (j - 1));
}
// ref UpsampleGrid :281-323: thread per (new node, channel)
__global__ void upsample_grid_kernel(int old_gs, const float *__restrict__ old_planes, float *__restrict__ new_planes)
{
  const int gs = old_gs * 2;
  const size_t Nn = (size_t)gs * gs * gs, No = (size_t)old_gs * old_gs * old_gs;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Nn * kCell) return;
  const size_t node = t / kCell;
  const int j = (int)(t % kCell);
  const int xyz[3] = {(int)(node % gs), (int)((node / gs) % gs), (int)(node / ((size_t)gs * gs))};
  const float step = 1.0f / (float)gs;
  int nc[3], fc[3];
  float lf[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float coord = (0.0f + step * ((float)xyz[a] + 0.5f)) * (float)old_gs;  // :294
    nc[a] = max(0, min(old_gs - 1, (int)coord));
    fc[a] = max(0, min(old_gs - 1, (int)coord + 1));
    lf[a] = coord - (float)nc[a];
  }
  auto at = [&](int x, int y, int z) { return plane_get(old_planes, No, (size_t)x + (size_t)y * old_gs + (size_t)z * old_gs * old_gs, j); };
  const float xy00 = lm_lerp(at(nc[0], nc[1], nc[2]), at(fc[0], nc[1], nc[2]), lf[0]);
  const float xy10 = lm_lerp(at(nc[0], fc[1], nc[2]), at(fc[0], fc[1], nc[2]), lf[0]);
  const float xy01 = lm_lerp(at(nc[0], nc[1], fc[2]), at(fc[0], nc[1], fc[2]), lf[0]);
  const float xy11 = lm_lerp(at(nc[0], fc[1], fc[2]), at(fc[0], fc[1], fc[2]), lf[0]);
  float v = lm_lerp(lm_lerp(xy00, xy10, lf[1]), lm_lerp(xy01, xy11, lf[1]), lf[2]);
  if (j == 0 && v < 0.0f) v = 0.0f;  // :316-317
  if (j == 0) new_planes[node] = v;
  else new_planes[Nn + node * 12 + (j - 1)] = v;
}

// un-relu'd trilinear density at the centre of lattice box (x,y,z): getDensity lambda :331-357 / RebuildOctree :459-483
__device__ __forceinline__ float box_centre_density(const SceneDev &S, int x, int y, int z)
{
  const int c[3] = {x, y, z};
  float lf[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float bmin = (float)c[a] * S.inv_gs_f, bmax = (float)(c[a] + 1) * S.inv_gs_f;
    const float p = (bmin + bmax) / 2.0f;
    lf[a] = (p - bmin) / (bmax - bmin);
  }
  SampleGeom g;
  g.lf[0] = lf[0]; g.lf[1] = lf[1]; g.lf[2] = lf[2];
  g.base = (uint32_t)x + (uint32_t)y * S.gs + (uint32_t)z * S.gs * S.gs;
  return density_at(S, g);
}
__constant__ int kChildOrder[8][3] = {{0, 0, 0}, {1, 1, 1}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}};  // :363-441
__global__ void box_child_flags_kernel(const __grid_constant__ SceneDev S, const int32_t *__restrict__ old_xyz, size_t n_old, uint32_t *__restrict__ flags)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_old * 8) return;
  const size_t b = t >> 3;
  const int c = (int)(t & 7);
  const int x = 2 * old_xyz[3 * b] + kChildOrder[c][0], y = 2 * old_xyz[3 * b + 1] + kChildOrder[c][1], z = 2 * old_xyz[3 * b + 2] + kChildOrder[c][2];
  flags[t] = box_centre_density(S, x, y, z) > 0.0f ? 1u : 0u;  // :369
}
__global__ void box_keep_flags_kernel(const __grid_constant__ SceneDev S, const int32_t *__restrict__ xyz, size_t n, uint32_t *__restrict__ flags)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  flags[t] = box_centre_density(S, xyz[3 * t], xyz[3 * t + 1], xyz[3 * t + 2]) > 0.0f ? 1u : 0u;  // :483
}
__global__ void box_compact_children_kernel(const int32_t *__restrict__ old_xyz, size_t n_old, const uint32_t *__restrict__ flags,
                                            const uint32_t *__restrict__ offs, int32_t *__restrict__ new_xyz)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_old * 8 || !flags[t]) return;
  const size_t b = t >> 3;
  const int c = (int)(t & 7);
  const size_t o = offs[t];
  new_xyz[3 * o] = 2 * old_xyz[3 * b] + kChildOrder[c][0];
  new_xyz[3 * o + 1] = 2 * old_xyz[3 * b + 1] + kChildOrder[c][1];
  new_xyz[3 * o + 2] = 2 * old_xyz[3 * b + 2] + kChildOrder[c][2];
}
__global__ void box_compact_kernel(const int32_t *__restrict__ xyz, size_t n, const uint32_t *__restrict__ flags, const uint32_t *__restrict__ offs,
                                   int32_t *__restrict__ new_xyz)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n || !flags[t]) return;
  const size_t o = offs[t];
  new_xyz[3 * o] = xyz[3 * t]; new_xyz[3 * o + 1] = xyz[3 * t + 1]; new_xyz[3 * o + 2] = xyz[3 * t + 2];
}
__global__ void build_box_maps_kernel(int nv, const int32_t *__restrict__ xyz, size_t n, uint32_t *__restrict__ occ, uint32_t *__restrict__ vox2id,
                                      uint32_t *__restrict__ occ8, int nb8)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t x = (uint32_t)xyz[3 * t], y = (uint32_t)xyz[3 * t + 1], z = (uint32_t)xyz[3 * t + 2];
  const uint32_t lin = x + y * nv + z * nv * nv;
  atomicOr(occ + (lin >> 5), 1u << (lin & 31));
  const uint32_t b = (x >> 3) + (y >> 3) * nb8 + (z >> 3) * nb8 * nb8;
  if (!((occ8[b >> 5] >> (b & 31)) & 1u)) atomicOr(occ8 + (b >> 5), 1u << (b & 31));
  vox2id[lin] = (uint32_t)t;
}
// brick distance field for the empty-space jump (FastMarch::try_jump): dist[b] = Chebyshev distance in bricks to the nearest occupied
// brick, capped at kBrickDistCap.  One init pass + kBrickDistCap relaxation passes over the nb8^3 bricks (ping-pong).
constexpr int kBrickDistCap = 40;
__global__ void brick_dist_init_kernel(int nb8, const uint32_t *__restrict__ occ8, uint8_t *__restrict__ dist)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= (uint32_t)nb8 * nb8 * nb8) return;
  dist[b] = ((occ8[b >> 5] >> (b & 31)) & 1u) ? 0 : kBrickDistCap;
}
__global__ void brick_dist_relax_kernel(int nb8, const uint8_t *__restrict__ in, uint8_t *__restrict__ out)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= (uint32_t)nb8 * nb8 * nb8) return;
  const int bx = b % nb8, by = (b / nb8) % nb8, bz = b / (nb8 * nb8);
  int m = in[b];
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        const int x = bx + dx, y = by + dy, z = bz + dz;
        if ((unsigned)x >= (unsigned)nb8 || (unsigned)y >= (unsigned)nb8 || (unsigned)z >= (unsigned)nb8) continue;   // nothing exists outside the lattice
        m = min(m, (int)in[(uint32_t)x + (uint32_t)y * nb8 + (uint32_t)z * nb8 * nb8] + 1);
      }
  out[b] = (uint8_t)m;
}
__global__ void count_nonzero_kernel(const uint8_t *__restrict__ marks, size_t n, unsigned long long *out)
{
  unsigned long long local = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) local += marks[i] ? 1 : 0;
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// two-level exclusive scan of uint32 flags (order-preserving box compaction)
constexpr int kScanBlock = 1024;
__global__ void scan_block_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n, uint32_t *__restrict__ block_sums)
{
  __shared__ uint32_t warp_sums[32];
  const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
  const uint32_t v = i < n ? in[i] : 0u;
  uint32_t x = v;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t s = warp_sums[lane];
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    warp_sums[lane] = s;
  }
  __syncthreads();
  const uint32_t incl = x + (wid ? warp_sums[wid - 1] : 0u);
  if (i < n) out[i] = incl - v;
  if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = incl;
}
__global__ void scan_sums_kernel(uint32_t *block_sums, size_t nb)  // single block, serial over chunks
{
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t base = 0; base < nb; base += kScanBlock) {
    const size_t i = base + threadIdx.x;
    const uint32_t v = i < nb ? block_sums[i] : 0u;
    uint32_t x = v;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t s = warp_sums[lane];
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const uint32_t incl = x + (wid ? warp_sums[wid - 1] : 0u) + carry;
    if (i < nb) block_sums[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry = incl;
    __syncthreads();
  }
}
__global__ void scan_add_kernel(uint32_t *__restrict__ out, size_t n, const uint32_t *__restrict__ block_sums)
{
  const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n) out[i] += block_sums[blockIdx.x];
}

inline unsigned blocks_for(size_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

// ------------------------------------------------------------------------------------------------
void launch_render_image(const SceneDev &S, uint32_t *out_dev, uint32_t w, uint32_t h, int grey, uint32_t *fb_count, uint32_t *fb_list,
                         cudaStream_t st, uint32_t y0, uint32_t y1, int lean, uint32_t *tap_n, float *tap_thr)
{
  cudaMemsetAsync(fb_count, 0, sizeof(uint32_t), st);
  if (y1 > h) y1 = h;
  if (y0 >= y1) return;
  const uint32_t hb = y1 - y0;
  const dim3 g(tile_grid(w, hb));
  if (lean && !S.occ && !grey && !S.rec) {
    const dim3 gl(LRF_LEAN_TILE2D ? ((w + kLeanBW - 1) / kLeanBW) * ((hb + kLeanBH - 1) / kLeanBH) : g.x);
    // The step-length rule (ref :145-146: a sample step longer than sqrt(3) * step is shortened) cannot fire between CONSECUTIVE face-adjacent
    // cells of a dense lattice (the step is at most half the diagonal of a 2 x 1 x 1 box = 1.2247 / gs < 1.7320 / gs), and compiling it out
    // of the loop is worth 2 % (2.05 -> 2.01 ms per view).  But behind a tie event tPrev can be the ENTRY point of a cell (a grazing box
    // listed after it), and the step then reaches the full diagonal: the rule stays in.  (Handling those samples in a separate cold
    // iteration keeps it exact but costs more registers than the branch: 2.13 ms.)
    const bool gate = true;
    if (tap_n) {
      if (gate) render_lean_kernel<true, true><<<gl, LRF_LEAN_BLOCK, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, y0, hb, tap_n, tap_thr);
      else render_lean_kernel<true, false><<<gl, LRF_LEAN_BLOCK, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, y0, hb, tap_n, tap_thr);
    } else {
      if (gate) render_lean_kernel<false, true><<<gl, LRF_LEAN_BLOCK, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, y0, hb, nullptr, nullptr);
      else render_lean_kernel<false, false><<<gl, LRF_LEAN_BLOCK, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, y0, hb, nullptr, nullptr);
    }
  } else if (S.occ) {
    if (grey) render_image_kernel<true, true><<<g, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, y0, hb);
    else render_image_kernel<true, false><<<g, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, y0, hb);
  } else {
    if (grey) render_image_kernel<false, true><<<g, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, y0, hb);
    else if (S.rec) render_image_kernel<false, false, true><<<g, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, y0, hb);
    else render_image_kernel<false, false><<<g, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, y0, hb);
  }
  render_fallback_kernel<<<148, kBlock, 0, st>>>(S, out_dev, w, h, grey, fb_count, fb_list);
}
// n views in one launch of the forward kernel launch_render_image would pick: frames w * h apart in out_dev, fb_list sized for n * w * h entries
void launch_render_views(const SceneDev &S, const float *views_dev, uint32_t n_views, uint32_t *out_dev, uint32_t w, uint32_t h, int grey, uint32_t *fb_count,
                         uint32_t *fb_list, cudaStream_t st, int lean)
{
  cudaMemsetAsync(fb_count, 0, sizeof(uint32_t), st);
  if (lean && !S.occ && !grey && !S.rec) {
    const uint32_t tiles = ((w + kLeanBW - 1) / kLeanBW) * ((h + kLeanBH - 1) / kLeanBH);
    render_lean_kernel<false, true, true><<<tiles * n_views, LRF_LEAN_BLOCK, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, 0, h, nullptr, nullptr, views_dev, tiles);
  } else {
    const uint32_t bpv = tile_grid(w, h);
    const dim3 g(bpv * n_views);
    if (S.occ) {
      if (grey) render_image_kernel<true, true><<<g, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, 0, h, views_dev, bpv);
      else render_image_kernel<true, false><<<g, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, 0, h, views_dev, bpv);
    } else {
      if (grey) render_image_kernel<false, true><<<g, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, 0, h, views_dev, bpv);
      else if (S.rec) render_image_kernel<false, false, true><<<g, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, 0, h, views_dev, bpv);
      else render_image_kernel<false, false><<<g, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list, 0, h, views_dev, bpv);
    }
  }
  render_fallback_kernel<<<148, kBlock, 0, st>>>(S, out_dev, w, h, grey, fb_count, fb_list, views_dev);
}
void launch_bake_records(const SceneDev &S, float4 *rec, float *max_abs_density, cudaStream_t st)
{
  const size_t N = (size_t)S.gs * S.gs * S.gs;
  cudaMemsetAsync(max_abs_density, 0, sizeof(float), st);
  bake_records_kernel<<<148 * 8, 256, 0, st>>>(S.dens, S.col, rec, N, max_abs_density);
}
size_t staged_smem_bytes(const StagedArgs &T) { return 1024 + kStCtl + (size_t)kStR * T.slot_bytes; }
int launch_render_staged(const SceneDev &S, const StagedArgs &T, int nw, uint32_t *out_dev, uint32_t w, uint32_t h, uint32_t *fb_count,
                         uint32_t *fb_list, uint32_t *tap_shaded, cudaStream_t st)
{
  const size_t smem = staged_smem_bytes(T);
  cudaMemsetAsync(fb_count, 0, sizeof(uint32_t), st);
  const uint32_t tiles_x = (w + 15) / 16;
  cudaError_t e;
  if (nw == 8) {
    e = cudaFuncSetAttribute(render_staged_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    render_staged_kernel<8><<<tiles_x * ((h + 15) / 16), 9 * 32, smem, st>>>(S, T, out_dev, w, h, fb_count, fb_list, tap_shaded);
  } else {
    e = cudaFuncSetAttribute(render_staged_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    render_staged_kernel<4><<<tiles_x * ((h + 7) / 8), 5 * 32, smem, st>>>(S, T, out_dev, w, h, fb_count, fb_list, tap_shaded);
  }
  render_list_kernel<<<148 * 2, kBlock, 0, st>>>(S, out_dev, w, h, fb_count, fb_list);
  return 0;
}
void launch_debug_image(const SceneDev &S, uint32_t w, uint32_t h, int grey, const DebugTaps &taps, int force_generic, cudaStream_t st)
{
  debug_kernel<true><<<tile_grid(w, h), kBlock, 0, st>>>(S, w, h, 0, nullptr, nullptr, grey, taps, force_generic);
}
void launch_debug_rays(const SceneDev &S, size_t n, const float *o, const float *d, int grey, const DebugTaps &taps, int force_generic,
                       cudaStream_t st)
{
  if (n == 0) return;
  debug_kernel<false><<<blocks_for(n, kBlock), kBlock, 0, st>>>(S, 0, 0, n, o, d, grey, taps, force_generic);
}
void launch_grad_image(const SceneDev &S, uint32_t w, uint32_t h, const float *ref_rgba, int grey, uint32_t *fb_count, uint32_t *fb_list,
                       cudaStream_t st)
{
  if (fb_count == nullptr) {   // generic thread-per-ray kernel for every pixel (kept for comparison: option "generic_image_grad")
    grad_kernel<true><<<tile_grid(w, h), kBlock, 0, st>>>(S, w, h, 0, nullptr, nullptr, ref_rgba, grey, nullptr, nullptr);
    return;
  }
  cudaMemsetAsync(fb_count, 0, sizeof(uint32_t), st);
  const dim3 g(tile_grid(w, h));
  if (S.occ) {
    if (grey) grad_image_fast_kernel<true, true><<<g, kBlock, 0, st>>>(S, w, h, ref_rgba, fb_count, fb_list);
    else grad_image_fast_kernel<true, false><<<g, kBlock, 0, st>>>(S, w, h, ref_rgba, fb_count, fb_list);
  } else {
    if (grey) grad_image_fast_kernel<false, true><<<g, kBlock, 0, st>>>(S, w, h, ref_rgba, fb_count, fb_list);
    else grad_image_fast_kernel<false, false><<<g, kBlock, 0, st>>>(S, w, h, ref_rgba, fb_count, fb_list);
  }
  // axis-aligned rays (at most a centre row / column): the generic march over the queue, usually an empty launch
  grad_kernel<true><<<blocks_for((size_t)w + h + 64, kBlock), kBlock, 0, st>>>(S, w, h, 0, nullptr, nullptr, ref_rgba, grey, fb_list, fb_count);
}
size_t grad_rays_hit_bytes(size_t n, int gs, uint32_t *cap_out)
{
  const uint32_t cap = 3u * (uint32_t)gs + 40u;   // a ray crosses < 3 gs cells; tie events add a few
  if (cap_out) *cap_out = cap;
  return n * cap * sizeof(HitRec);
}
void launch_grad_rays(const SceneDev &S, size_t n, const float *o, const float *d, const float *ref_rgba, int grey, uint32_t *fb_count,
                      uint32_t *fb_list, void *hit_scratch, cudaStream_t st, int parallel_gather)
{
  if (n == 0) return;
  cudaMemsetAsync(fb_count, 0, sizeof(uint32_t), st);
  const unsigned g = blocks_for(n * 32, kWrayBlock);
  uint32_t cap = 0;
  grad_rays_hit_bytes(n, S.gs, &cap);
  HitRec *hb = (HitRec *)hit_scratch;
  if (S.occ) {
    if (grey) grad_rays_warp_kernel<true, true><<<g, kWrayBlock, 0, st>>>(S, n, o, d, ref_rgba, fb_count, fb_list, hb, cap, parallel_gather);
    else grad_rays_warp_kernel<true, false><<<g, kWrayBlock, 0, st>>>(S, n, o, d, ref_rgba, fb_count, fb_list, hb, cap, parallel_gather);
  } else {
    if (grey) grad_rays_warp_kernel<false, true><<<g, kWrayBlock, 0, st>>>(S, n, o, d, ref_rgba, fb_count, fb_list, hb, cap, parallel_gather);
    else grad_rays_warp_kernel<false, false><<<g, kWrayBlock, 0, st>>>(S, n, o, d, ref_rgba, fb_count, fb_list, hb, cap, parallel_gather);
  }
  // rays the warp kernel cannot take (axis-aligned): the generic thread-per-ray march, usually an empty launch
  grad_kernel<false><<<blocks_for(n, kBlock), kBlock, 0, st>>>(S, 0, 0, n, o, d, ref_rgba, grey, fb_list, fb_count);
}
void launch_gen_rays(const float *proj_inv_host16, size_t n, const float *view_inv_dev, const uint32_t *picks, uint32_t w, uint32_t h, float *o,
                     float *d, cudaStream_t st)
{
  if (n == 0) return;
  SceneDev S{};
  for (int i = 0; i < 16; i++) S.proj_inv[i] = proj_inv_host16[i];
  gen_rays_kernel<<<blocks_for(n, 256), 256, 0, st>>>(n, view_inv_dev, picks, w, h, o, d, S);
}
void launch_tv_sweep(const CUtensorMap &tm, int gs, float *gdens, int z_begin, int z_end, int atomic, uint32_t *diag, cudaStream_t st, int blocks_per_sm)
{
  if (z_end <= z_begin) return;
  const int chunks = (z_end - z_begin + kSwZChunk - 1) / kSwZChunk;
  const int n_work = (gs / kSwX) * (gs / kSwY) * chunks;
  // blocks_per_sm > 0: the kernel runs beside an HBM-bound one and must leave it room on every SM: a persistent grid of that many blocks
  const int grid = blocks_per_sm > 0 ? std::min(n_work, 148 * blocks_per_sm) : n_work;
  tv_sweep_kernel<<<grid, 256, 0, st>>>(tm, gs, gdens, z_begin, z_end, atomic, diag);
}
void launch_tv_grad(int gs, const float *dens, float *gdens, float *tmp_v, float *tmp_g, cudaStream_t st)
{
  const size_t N = (size_t)gs * gs * gs;
  if (gs >= kTvT) {
    const int tiles = gs / kTvT;
    tv_fused_kernel<<<tiles * tiles * tiles, 256, 0, st>>>(gs, dens, gdens, 0, 0, tiles * tiles * tiles);
    return;
  }
  tv_voxel_kernel<<<blocks_for(N, 256), 256, 0, st>>>(gs, dens, tmp_v);
  tv_gather_kernel<<<blocks_for(N, 256), 256, 0, st>>>(gs, tmp_v, tmp_g);
  tv_scatter_kernel<<<blocks_for(N, 256), 256, 0, st>>>(gs, tmp_g, gdens);
}
void launch_adam(float *p, float *g, float *V, float *S, size_t off, size_t count, float lr, float beta1, float beta2, float eps, double om_b1,
                 double om_b2, int zero_grad, cudaStream_t st)
{
  if (count == 0) return;
  AdamConst C;
  C.lr = lr; C.beta1 = beta1; C.beta2 = beta2; C.eps = eps;
  C.omb1f = 1 - beta1; C.omb2f = 1 - beta2;   // (1 - beta) evaluated in float, as the reference's int - float
  C.om_b1 = om_b1; C.om_b2 = om_b2; C.inv_om_b1 = 1.0 / om_b1; C.inv_om_b2 = 1.0 / om_b2;
  adam_kernel<<<blocks_for((count + 3) / 4, 256), 256, 0, st>>>(p + off, g + off, V + off, S + off, count, C, zero_grad);
}
void launch_tv_grad_slab(int gs, const float *dens, float *gdens, int z_begin, int z_end, int atomic, int max_blocks, cudaStream_t st)
{
  const int tiles = gs / kTvT, l0 = z_begin / kTvT, l1 = z_end / kTvT;
  const int n_tiles = tiles * tiles * (l1 - l0);
  if (n_tiles <= 0) return;
  const int grid = (max_blocks > 0 && max_blocks < n_tiles) ? max_blocks : n_tiles;
  tv_fused_kernel<<<grid, 256, 0, st>>>(gs, dens, gdens, l0, atomic, n_tiles);
}
void launch_adam_peer(int world, int rank, float *const *grads, float *const *params, float *V, float *S, size_t off, size_t count, float lr,
                      float beta1, float beta2, float eps, double om_b1, double om_b2, cudaStream_t st)
{
  if (count == 0) return;
  AdamConst C;
  C.lr = lr; C.beta1 = beta1; C.beta2 = beta2; C.eps = eps;
  C.omb1f = 1 - beta1; C.omb2f = 1 - beta2;
  C.om_b1 = om_b1; C.om_b2 = om_b2; C.inv_om_b1 = 1.0 / om_b1; C.inv_om_b2 = 1.0 / om_b2;
  PeerSet P{};
  P.world = world; P.rank = rank;
  for (int r = 0; r < world; r++) { P.grad[r] = grads[r]; P.param[r] = params[r]; }
  adam_peer_kernel<<<blocks_for((count + 3) / 4, 256), 256, 0, st>>>(P, V, S, off, count, C);
}
void launch_adam_mc(const float *mc_grad, float *mc_param, const float *param, float *V, float *S, size_t off, size_t count, float lr, float beta1,
                    float beta2, float eps, double om_b1, double om_b2, cudaStream_t st, int unroll, int blocks_per_sm)
{
  if (count == 0) return;
  AdamConst C;
  C.lr = lr; C.beta1 = beta1; C.beta2 = beta2; C.eps = eps;
  C.omb1f = 1 - beta1; C.omb2f = 1 - beta2;
  C.om_b1 = om_b1; C.om_b2 = om_b2; C.inv_om_b1 = 1.0 / om_b1; C.inv_om_b2 = 1.0 / om_b2;
  if (unroll != 2 && unroll != 8) unroll = 4;
  unsigned grid = blocks_for((count + 3) / 4, 256 * unroll);
  if (blocks_per_sm > 0 && grid > 148u * blocks_per_sm) grid = 148u * blocks_per_sm;
  if (unroll == 2) adam_mc_kernel<2><<<grid, 256, 0, st>>>(mc_grad, mc_param, param, V, S, off, count, C);
  else if (unroll == 8) adam_mc_kernel<8><<<grid, 256, 0, st>>>(mc_grad, mc_param, param, V, S, off, count, C);
  else adam_mc_kernel<4><<<grid, 256, 0, st>>>(mc_grad, mc_param, param, V, S, off, count, C);
}
void launch_peer_barrier(int world, int rank, uint32_t *const *flags, uint32_t epoch, uint32_t *timed_out, cudaStream_t st)
{
  PeerFlags F{};
  F.world = world; F.rank = rank;
  for (int r = 0; r < world; r++) F.flags[r] = flags[r];
  peer_barrier_kernel<<<1, 32, 0, st>>>(F, epoch, timed_out);
}
void launch_fill(float *p, size_t count, float v, cudaStream_t st)
{
  if (count == 0) return;
  fill_kernel<<<blocks_for((count + 3) / 4, 256), 256, 0, st>>>(p, count, v);
}
void launch_aos_to_planes(const float *aos, float *planes, size_t N, cudaStream_t st)
{
  aos_to_planes_kernel<<<blocks_for(N * kCell, 256), 256, 0, st>>>(aos, planes, N);
}
void launch_planes_to_aos(const float *planes, float *aos, size_t N, cudaStream_t st)
{
  planes_to_aos_kernel<<<blocks_for(N * kCell, 256), 256, 0, st>>>(planes, aos, N);
}
void launch_upsample_grid(int old_gs, const float *old_planes, float *new_planes, cudaStream_t st)
{
  const size_t Nn = (size_t)old_gs * old_gs * old_gs * 8;
  upsample_grid_kernel<<<blocks_for(Nn * kCell, 256), 256, 0, st>>>(old_gs, old_planes, new_planes);
}
void launch_box_child_flags(const SceneDev &S_new, const int32_t *old_xyz, size_t n_old, uint32_t *flags, cudaStream_t st)
{
  if (n_old == 0) return;
  box_child_flags_kernel<<<blocks_for(n_old * 8, 256), 256, 0, st>>>(S_new, old_xyz, n_old, flags);
}
void launch_box_keep_flags(const SceneDev &S, const int32_t *xyz, size_t n, uint32_t *flags, cudaStream_t st)
{
  if (n == 0) return;
  box_keep_flags_kernel<<<blocks_for(n, 256), 256, 0, st>>>(S, xyz, n, flags);
}
void launch_box_compact_children(const int32_t *old_xyz, size_t n_old, const uint32_t *flags, const uint32_t *offsets, int32_t *new_xyz,
                                 cudaStream_t st)
{
  if (n_old == 0) return;
  box_compact_children_kernel<<<blocks_for(n_old * 8, 256), 256, 0, st>>>(old_xyz, n_old, flags, offsets, new_xyz);
}
void launch_box_compact(const int32_t *xyz, size_t n, const uint32_t *flags, const uint32_t *offsets, int32_t *new_xyz, cudaStream_t st)
{
  if (n == 0) return;
  box_compact_kernel<<<blocks_for(n, 256), 256, 0, st>>>(xyz, n, flags, offsets, new_xyz);
}
void launch_build_box_maps(int nv, const int32_t *xyz, size_t n, uint32_t *occ, uint32_t *vox2id, uint32_t *occ8, int nb8, cudaStream_t st)
{
  if (n == 0) return;
  build_box_maps_kernel<<<blocks_for(n, 256), 256, 0, st>>>(nv, xyz, n, occ, vox2id, occ8, nb8);
}
void launch_brick_dist(int nb8, const uint32_t *occ8, uint8_t *dist, uint8_t *tmp, cudaStream_t st)
{
  const unsigned g = blocks_for((size_t)nb8 * nb8 * nb8, 256);
  brick_dist_init_kernel<<<g, 256, 0, st>>>(nb8, occ8, dist);
  uint8_t *a = dist, *b = tmp;
  for (int it = 0; it < kBrickDistCap; it++) {   // an even number of passes: the result ends in `dist`
    brick_dist_relax_kernel<<<g, 256, 0, st>>>(nb8, a, b);
    uint8_t *t = a; a = b; b = t;
  }
  static_assert(kBrickDistCap % 2 == 0, "ping-pong parity");
}
void launch_count_nonzero(const uint8_t *marks, size_t n, unsigned long long *out, cudaStream_t st)
{
  count_nonzero_kernel<<<148 * 8, 256, 0, st>>>(marks, n, out);
}
void launch_ingest_rgba8(const uint8_t *src_rgba8, float *dst_rgba, uint32_t w, uint32_t h, int flip, const float *lut_dev, cudaStream_t st)
{
  const size_t n = (size_t)w * h;
  if (n == 0) return;
  const unsigned blocks = (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  ingest_rgba8_kernel<<<blocks, 256, 0, st>>>((const uchar4 *)src_rgba8, (float4 *)dst_rgba, w, h, flip, lut_dev);
}
void launch_gather_targets(size_t n, const uint32_t *picks, const float *targets, size_t view_stride_px, float *out_rgba, cudaStream_t st)
{
  if (n == 0) return;
  gather_targets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, picks, (const float4 *)targets, view_stride_px, (float4 *)out_rgba);
}
void launch_pack_bmp(const uint32_t *pix, uint8_t *out, size_t n, cudaStream_t st)
{
  if (n == 0) return;
  const unsigned blocks = (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  pack_bmp_kernel<<<blocks, 256, 0, st>>>(pix, out, n);
}
void launch_fill_synthetic(int gs, float D, int mode, float *planes, cudaStream_t st)
{
  const size_t N = (size_t)gs * gs * gs;
  fill_synthetic_kernel<<<blocks_for(N, 256), 256, 0, st>>>(gs, D, mode, planes, (float4 *)(planes + N));
}
void launch_dense_boxes(int n_vox_axis, int32_t *xyz, cudaStream_t st)
{
  const size_t total = (size_t)n_vox_axis * n_vox_axis * n_vox_axis;
  if (total) dense_boxes_kernel<<<blocks_for(total, 256), 256, 0, st>>>(n_vox_axis, xyz);
}
size_t scan_tmp_elems(size_t n) { return (n + kScanBlock - 1) / kScanBlock + 1; }
void launch_exclusive_scan(const uint32_t *in, uint32_t *out, size_t n, uint32_t *block_tmp, cudaStream_t st)
{
  if (n == 0) return;
  const size_t nb = (n + kScanBlock - 1) / kScanBlock;
  scan_block_kernel<<<(unsigned)nb, kScanBlock, 0, st>>>(in, out, n, block_tmp);
  scan_sums_kernel<<<1, kScanBlock, 0, st>>>(block_tmp, nb);
  scan_add_kernel<<<(unsigned)nb, kScanBlock, 0, st>>>(out, n, block_tmp);
}

}  // namespace lrf
