// lrf_staged.cuh — forward render of dense grids, re-decomposed for sm_100a (round 2).
//
// ref: kernel2D_RayMarch example_tracer.cpp:627-663 -> RayGridIntersection :120-194 (the ~60-line sample loop).
//
// Why a second forward kernel.  ncu on the thread-per-ray gather kernel (render_image_kernel, profiles/r1t_*) shows the L1
// data pipe 88 % busy: 147 LSU wavefronts per warp-iteration, of which ~100 are 12 LDG.128 per lane that each touch 1-3
// 128-byte lines per quarter-warp, and 48 are the shared-memory colour window.  The 32 rays of a warp read ~6.6 KB per step
// out of ~1 KB of distinct node data.  This kernel removes both costs:
//
//  * BAKED NODE RECORDS.  The render reads a second copy of the parameters as one 64-byte record per node
//    {density, R, G, B | sh0..sh3 | sh4..sh7 | sh8, 0, 0, 0} (bake_records_kernel; rebuilt lazily when the parameters change).
//  * TMA-STAGED PLANES.  A CTA owns a 16x8 (or 16x16) pixel tile and marches it along the tile's dominant axis A.  A
//    producer warp streams, per node plane, the tile's footprint rectangle (W x H records) from HBM/L2 into a ring of
//    kStR shared-memory slots with ONE cp.async.bulk.tensor (4-D tensor map {16 floats, x, y, z}; the box is 1 node thick
//    along A) signalled through an mbarrier — no LSU wavefronts, no registers, zero fill outside the grid.  The box origin
//    per plane comes from the tile's four corner rays (a pixel rectangle maps to a convex quadrilateral on every plane).
//  * The consumer lanes run the exact per-ray DDA free of each other (no slab lock-step: a lane only needs the planes of
//    its voxel resident; the ring slides with the slowest lane of the CTA) and gather their nodes from shared memory with
//    LDS.128 (broadcast for lanes in the same voxel).
//  * REGISTER WINDOW.  The 2x2x2 window {density, R, G, B, q = sh . Y(d)} lives in registers in parity slots (as in
//    lrf_fast.cuh) but is committed by predicated loads / predicated final FMAs straight into the slot registers, and the
//    slot addresses are carried incrementally (a face step adds one delta to the 4 stale slots).
//  * FILTERED DECISIONS.  What the parity contract fixes bit-exactly is the hit list and the shaded-sample count, i.e. the
//    DDA order and the `throughput < 0.1` break (ref :131).  The DDA keeps the fp32-filter + exact-fp64 comparison of
//    lrf_fast.cuh.  The break is decided on an APPROXIMATE throughput (fused weight-form density interpolation, ex2.approx)
//    with a running error bound: a ray whose throughput ever falls inside [0.1 - band, 0.1 + band) is not decided here — it
//    is appended to the fallback list and rendered by the exact kernel (FastMarch, libm-exact expf).  Everything that would
//    need the tie machinery (edge / corner grazing, entry on a plane, axis-aligned rays) goes to the same list, so the hot
//    loop carries no queue / resync state.  Sample geometry (t, delta, fractional position) is op-for-op the reference's.
//
// Results: identical shaded counts (tests/test_gpu_staged.py through the tap below), RGB within the 1e-4 tolerance.
#pragma once
#include "lrf_kernels.cuh"
#include "lrf_fast.cuh"

#ifndef LRF_ST_SWZ
#define LRF_ST_SWZ 0          // 1: CU_TENSOR_MAP_SWIZZLE_64B on the staged planes (records 2 apart in a row stop sharing banks)
#endif
#ifndef LRF_ST_DEBUG
#define LRF_ST_DEBUG 0        // 1: every consumer warp leaves its ring state in shared memory; the producer's watchdog copies it out
#endif
#ifndef LRF_ST_RELOAD
#define LRF_ST_RELOAD 1       // 1: {density, R, G, B} of all 8 corners re-read from the staged planes every sample (window = q only: 32 registers less);
#endif                        // 0: kept in the register window, committed by predicated loads
#ifndef LRF_ST_INFLIGHT
#define LRF_ST_INFLIGHT 8     // planes the producer keeps in flight (<= ring size)
#endif
#ifndef LRF_ST_MINB
#define LRF_ST_MINB 3         // resident CTAs per SM the register allocation is sized for
#endif

namespace lrf {

constexpr int kStR = 8;                       // ring slots (power of two)
constexpr int kStRecF = 16;                   // floats per baked record
constexpr int kStRecB = kStRecF * 4;
constexpr uint32_t kStCtl = 512;              // control block in front of the ring: mbarriers, plane bases, progress words
constexpr uint32_t kStSlotAlign = 512;        // ring slots start on the swizzle period
constexpr int kStDone = 0x7fffffff;
constexpr float kStMargin = 0.03125f;         // voxels of slack around the corner-ray footprint (rounding of its fp32 evaluation)
// error budget of the approximate throughput, relative, per shaded sample: ex2.approx (2 ulp) + the rounding of its argument
// + libm's own 0.51 ulp + the two roundings of the running products
constexpr float kStErrStep = 8.0e-7f;
constexpr float kStErrX = 0x1p-21f;           // ... + |sigma * delta| * 2^-21 (argument roundings grow with the argument)


__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)   // non-blocking
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)   // suspends the warp in hardware for a bounded time
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Watchdog: a wait that lasts longer than ~2 s is a protocol bug or a lost TMA, never a slow one.  The lane leaves a record in
// host-mapped memory (it survives the trap) and kills the kernel instead of hanging the GPU.
constexpr long long kStWatchdogCycles = 4000000000LL;
__device__ __noinline__ void staged_die(uint32_t *diag, uint32_t code, uint32_t a, uint32_t b, uint32_t c)
{
  if (diag) {
    if (atomicCAS(diag, 0u, code) == 0u) {
      diag[1] = blockIdx.x; diag[2] = threadIdx.x; diag[3] = a; diag[4] = b; diag[5] = c; diag[6] = __activemask();
      __threadfence_system();
    }
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t *diag, uint32_t code, uint32_t a, uint32_t b)
{
  // Called by whole warps only.  ONE lane watches the barrier and the warp meets at an explicit warp barrier: a per-lane wait loop
  // lets the lanes leave at different times, nothing reconverges them afterwards, and the code that follows (convergent warp
  // reductions, the elected TMA issue) was then run by partial warps on their own (profiles/r2_staged_debug.txt).
  if ((threadIdx.x & 31) == 0 && !mbar_try(bar, parity)) {
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
      if (clock64() - t0 > kStWatchdogCycles) staged_die(diag, code, a, b, parity);
  }
  __syncwarp();
}
// the same for ONE calling thread (the caller publishes the result with a block / warp barrier)
__device__ __forceinline__ void mbar_wait_one(uint32_t bar, uint32_t parity, uint32_t *diag, uint32_t code, uint32_t a, uint32_t b)
{
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > kStWatchdogCycles) staged_die(diag, code, a, b, parity);
}
// non-blocking, same answer in every lane
__device__ __forceinline__ bool mbar_test_warp(uint32_t bar, uint32_t parity)
{
  uint32_t ok = 0;
  if ((threadIdx.x & 31) == 0) ok = mbar_test(bar, parity) ? 1u : 0u;
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2, int c3)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(tm),
               "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ int lds_acquire_s32(uint32_t a) { int v; asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_release_s32(uint32_t a, int v) { asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_f4(uint32_t a) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f1(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
// chunk 0 of a record {density, R, G, B} straight into the window slot that is stale: member A of the antipodal pair when toA, else member B
__device__ __forceinline__ void lds_slot(uint32_t a, uint32_t toA, float4 &A, float4 &B)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %8, 0;\n\t"
      "@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%9];\n\t"
      "@!p ld.shared.v4.f32 {%4, %5, %6, %7}, [%9];\n\t}"
      : "+f"(A.x), "+f"(A.y), "+f"(A.z), "+f"(A.w), "+f"(B.x), "+f"(B.y), "+f"(B.z), "+f"(B.w)
      : "r"(toA), "r"(a));
}
// same with one address register per member (what the march uses)
__device__ __forceinline__ void lds_slot2(uint32_t aA, uint32_t aB, uint32_t toA, float4 &A, float4 &B)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %8, 0;\n\t"
      "@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%9];\n\t"
      "@!p ld.shared.v4.f32 {%4, %5, %6, %7}, [%10];\n\t}"
      : "+f"(A.x), "+f"(A.y), "+f"(A.z), "+f"(A.w), "+f"(B.x), "+f"(B.y), "+f"(B.z), "+f"(B.w)
      : "r"(toA), "r"(aA), "r"(aB));
}
// rare paths of the march, kept out of line (they would otherwise be if-converted into the hot loop)
// exact comparison of the three exit-plane times, the reference's own fp64 quotients (ref octree.h:449-488): the crossing mask
static __device__ __noinline__ int staged_exact_cross(const float *pa, const float *o, const float *d)
{
  double te[3];
#pragma unroll
  for (int a = 0; a < 3; a++) te[a] = ((double)pa[a] - (double)o[a]) / (double)d[a];
  const double em = fmin(te[0], fmin(te[1], te[2]));
  return (te[0] == em ? 1 : 0) | (te[1] == em ? 2 : 0) | (te[2] == em ? 4 : 0);
}
// ref example_tracer.cpp:145-146
static __device__ __noinline__ float staged_long_step(float delta, double step)
{
  return (float)((double)delta - (double)((int)((double)delta / step) - 1) * step);
}
// q = fma(s8, y8, part) into the stale member of the pair
__device__ __forceinline__ void fma_slot(uint32_t toA, float s8, float y8, float part, float &qA, float &qB)
{
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p fma.rn.f32 %0, %3, %4, %5;\n\t@!p fma.rn.f32 %1, %3, %4, %5;\n\t}" : "+f"(qA), "+f"(qB) : "r"(toA), "f"(s8), "f"(y8), "f"(part));
}
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float pick3(const float *v, int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }
__device__ __forceinline__ int pick3i(const int *v, int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }
// address of 16-byte chunk 0 of the record at r, as the TMA stored it (chunk k is at the result ^ (k << 4) when swizzled, + 16 k otherwise)
__device__ __forceinline__ uint32_t rec_chunk0(uint32_t r)
{
#if LRF_ST_SWZ
  return r ^ ((r >> 3) & 0x30u);
#else
  return r;
#endif
}
__device__ __forceinline__ uint32_t rec_chunk(uint32_t c0, int k)
{
#if LRF_ST_SWZ
  return c0 ^ ((uint32_t)k << 4);
#else
  return c0 + 16u * (uint32_t)k;
#endif
}

// planes -> records; also the largest |density| (the error bound of the approximate interpolation scales with it)
__global__ void __launch_bounds__(256) bake_records_kernel(const float *__restrict__ dens, const float4 *__restrict__ col, float4 *__restrict__ rec, size_t N,
                                                           float *__restrict__ max_abs_density)
{
  float m = 0.0f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < N; i += (size_t)gridDim.x * 256) {
    const float d = __ldg(dens + i);
    const float4 a = __ldg(col + i * kColF4), b = __ldg(col + i * kColF4 + 1), e = __ldg(col + i * kColF4 + 2);
    float4 *r = rec + i * (kStRecF / 4);
    r[0] = make_float4(d, a.x, a.y, a.z);
    r[1] = make_float4(a.w, b.x, b.y, b.z);
    r[2] = make_float4(b.w, e.x, e.y, e.z);
    r[3] = make_float4(e.w, 0.0f, 0.0f, 0.0f);
    const float ad = fabsf(d);
    m = (ad > m || ad != ad) ? ad : m;                          // NaN sticks: every break then counts as ambiguous and the exact kernel decides
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, m, s);
    m = (o > m || o != o) ? o : m;
  }
  if ((threadIdx.x & 31) == 0) {
    if (m != m) atomicExch((unsigned int *)max_abs_density, 0x7fc00000u);
    else atomicMax((int *)max_abs_density, __float_as_int(m));   // non-negative floats order like their bit patterns; a NaN pattern stays the largest
  }
}

// exact first cell of a ray (same construction as FastMarch::run / the generic march): 0 = misses the lattice, 1 = cur[] is the
// first cell, 2 = the ray starts on a lattice plane shared by two cells (a tie event: the exact kernel handles it)
__device__ __forceinline__ int staged_entry(const SceneDev &S, const float *of, const float *df, int *cur)
{
  const int n = S.n;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  double t_enter = -INF, t_exit = INF, od[3], hd[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    od[a] = (double)of[a];
    hd[a] = (double)df[a];
    const double T0 = (0.0 - od[a]) / hd[a], Tn = ((double)n * S.inv_gs - od[a]) / hd[a];
    const bool pos = df[a] > 0.0f;
    t_enter = fmax(t_enter, pos ? T0 : Tn);
    t_exit = fmin(t_exit, pos ? Tn : T0);
  }
  const double t_start = fmax(t_enter, 0.0);
  if (!(t_exit >= t_start)) return 0;                            // ref octree.h:483 for the whole lattice
  int two = 0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const int s = df[a] > 0.0f ? 1 : -1;
    auto Tex = [&](int j) { return ((double)j * S.inv_gs - od[a]) / hd[a]; };
    int i = (int)floor((od[a] + t_start * hd[a]) * (double)S.gs);
    i = max(0, min(n - 1, i));
    for (int it = 0; it < 4; it++) {
      if (Tex(i + (s > 0 ? 0 : 1)) > t_start && (unsigned)(i - s) < (unsigned)n) { i -= s; continue; }
      if (Tex(i + (s > 0 ? 1 : 0)) < t_start && (unsigned)(i + s) < (unsigned)n) { i += s; continue; }
      break;
    }
    cur[a] = i;
    if (Tex(i + (s > 0 ? 0 : 1)) == t_start && (unsigned)(i - s) < (unsigned)n) two |= 1 << a;
    else if (Tex(i + (s > 0 ? 1 : 0)) == t_start && (unsigned)(i + s) < (unsigned)n) two |= 1 << a;
  }
  return two ? 2 : 1;
}

// NW consumer warps (2 wide: a 16 x 2 NW pixel tile of 8 x 4 warp tiles, quarter-warps on 4 x 2 pixel blocks) + 1 producer warp.
// tap_shaded: optional per-pixel shaded-sample count of the rays this kernel decided (0xffffffff for rays sent to the exact kernel)
template <int NW>
__global__ void __launch_bounds__((NW + 1) * 32, NW == 4 ? LRF_ST_MINB : (LRF_ST_MINB + 1) / 2)
render_staged_kernel(const __grid_constant__ SceneDev S, const __grid_constant__ StagedArgs T, uint32_t *__restrict__ out, uint32_t w, uint32_t h,
                     uint32_t *__restrict__ fb_count, uint32_t *__restrict__ fb_list, uint32_t *__restrict__ tap_shaded)
{
  constexpr int kTileW = 16, kTileH = 2 * NW;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);                        // 1024-aligned: dynamic shared memory starts the CTA's window
  if (sbase & 1023u) staged_die(T.diag, 5u, sbase, 0u, 0u);
  const uint32_t a_full = sbase, a_base = sbase + 64, a_first = sbase + 96, a_prog = sbase + 128, a_ovf = sbase + 160, a_ring = sbase + kStCtl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles_x = (w + kTileW - 1) / kTileW;
  const uint32_t tile_x0 = (blockIdx.x % tiles_x) * kTileW, tile_y0 = (blockIdx.x / tiles_x) * kTileH;
  const int gs = S.gs, n = S.n;

  // ---- march axis of the tile: the dominant axis of its centre ray; lateral axes U < V in memory order ----
  int A;
  bool a_pos;
  {
    float oc[3], dc[3];
    gen_ray(S.view_inv, S.proj_inv, min(tile_x0 + kTileW / 2, w - 1), min(tile_y0 + kTileH / 2, h - 1), w, h, oc, dc);
    const float ax = fabsf(dc[0]), ay = fabsf(dc[1]), az = fabsf(dc[2]);
    A = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    a_pos = pick3(dc, A) > 0.0f;
  }
  const int U = (A == 0) ? 1 : 0, V = (A == 2) ? 1 : 2;
  const int W = T.W[A], H = T.H[A];

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStR; s++) mbar_init(a_full + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  if (warp == NW) {
    // =========================== producer warp ===========================
    // lanes 0..3: the corner rays of the tile.  On node plane k (index along A) a ray sits at u(k) = cu + k * su (voxel units); plane k
    // serves the slabs on both sides of it, where the ray's lateral coordinate stays within u(k) +- |su|.
    float cu = 0.0f, su = 0.0f, cv = 0.0f, sv = 0.0f;
    bool corner_ok = true;
    if (lane < 4) {
      float o[3], d[3];
      const uint32_t cx = min(tile_x0 + ((lane & 1) ? kTileW - 1 : 0), w - 1), cy = min(tile_y0 + ((lane & 2) ? kTileH - 1 : 0), h - 1);
      gen_ray(S.view_inv, S.proj_inv, cx, cy, w, h, o, d);
      const float dA = pick3(d, A), oA = pick3(o, A) * S.gs_f;
      corner_ok = ((dA > 0.0f) == a_pos) && fabsf(dA) >= 0.25f;
      su = pick3(d, U) / dA; sv = pick3(d, V) / dA;
      cu = pick3(o, U) * S.gs_f - oA * su; cv = pick3(o, V) * S.gs_f - oA * sv;
    }
    auto box_of = [&](int k, int &u0, int &v0, int &u1, int &v1) {   // warp-uniform result; lanes >= 4 contribute neutral values
      const float fu = fmaf((float)k, su, cu), fv = fmaf((float)k, sv, cv);
      const bool c = lane < 4;
      const int lu = c ? (int)floorf(fu - fabsf(su) - kStMargin) : 0x7fffffff, hu = c ? (int)floorf(fu + fabsf(su) + kStMargin) + 1 : (int)0x80000000;
      const int lv = c ? (int)floorf(fv - fabsf(sv) - kStMargin) : 0x7fffffff, hv = c ? (int)floorf(fv + fabsf(sv) + kStMargin) + 1 : (int)0x80000000;
      u0 = __reduce_min_sync(FULL, lu); u1 = __reduce_max_sync(FULL, hu);
      v0 = __reduce_min_sync(FULL, lv); v1 = __reduce_max_sync(FULL, hv);
    };
    {
      // the lateral extent is a convex function of k: checking the two end planes covers every plane
      int u0, v0, u1, v1;
      bool fits = __all_sync(FULL, corner_ok);
      box_of(0, u0, v0, u1, v1);
      fits = fits && (u1 - u0 + 1 <= W) && (v1 - v0 + 1 <= H);
      box_of(gs - 1, u0, v0, u1, v1);
      fits = fits && (u1 - u0 + 1 <= W) && (v1 - v0 + 1 <= H);
      if (lane == 0) sts_s32(a_ovf, fits ? 0 : 1);
    }
    __syncthreads();
    if (lds_s32(a_ovf)) return;
    int p0 = kStDone;
#pragma unroll
    for (int i = 0; i < NW; i++) p0 = min(p0, lds_s32(a_first + 4 * i));
    if (p0 == kStDone) return;
    const int q_end = gs + 2 - p0;                                // planes p0 .. gs-1 are real; two virtual ones let the lanes of the last slab run
    int q = 0, qmin = 0;
    for (; q < q_end; q++) {
      // ring flow control: slot (q & R-1) is free once every consumer lane has left plane q - R behind
      long long t0 = 0;
      __syncwarp();
      while (q >= qmin + kStR) {
        int m = kStDone;
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < NW; i++) m = min(m, lds_acquire_s32(a_prog + 4 * i));
        }
        qmin = __shfl_sync(FULL, m, 0);   // ONE view of the progress words for the whole warp
        if (qmin == kStDone || q < qmin + kStR) break;
        __nanosleep(200);
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > kStWatchdogCycles) {
#if LRF_ST_DEBUG
          if (lane == 0 && T.diag && atomicCAS(T.diag + 7, 0u, 1u) == 0u) {   // one CTA's picture, not a mixture
            for (int i = 0; i < NW; i++) T.diag[8 + i] = (uint32_t)lds_acquire_s32(a_prog + 4 * i);
            for (int i = 0; i < NW * 8; i++) T.diag[16 + i] = (uint32_t)lds_acquire_s32(sbase + 192 + 4 * i);
            for (int i = 0; i < kStR; i++) T.diag[16 + NW * 8 + i] = (uint32_t)mbar_test(a_full + 8 * i, 0) | ((uint32_t)mbar_test(a_full + 8 * i, 1) << 1);
            for (int i = 0; i < 2 * kStR; i++) T.diag[24 + NW * 8 + i] = (uint32_t)lds_s32(sbase + 448 + 4 * i);
            T.diag[40 + NW * 8] = (uint32_t)W; T.diag[41 + NW * 8] = (uint32_t)H; T.diag[42 + NW * 8] = T.plane_bytes[A]; T.diag[43 + NW * 8] = T.slot_bytes;
            T.diag[44 + NW * 8] = (uint32_t)q; T.diag[45 + NW * 8] = (uint32_t)qmin; T.diag[46 + NW * 8] = blockIdx.x;
            __threadfence_system();
            __nanosleep(100000);
          }
#endif
          staged_die(T.diag, 1u, (uint32_t)q, (uint32_t)qmin, (uint32_t)p0);
        }
      }
      if (qmin == kStDone) break;
      const int p = p0 + q, k = a_pos ? p : gs - 1 - p, slot = q & (kStR - 1);
      const uint32_t bar = a_full + 8 * slot, dst = a_ring + (uint32_t)slot * T.slot_bytes;
      // at most LRF_ST_INFLIGHT planes in flight; with the default (= ring size) this is the producer's own check that the slot's
      // previous plane has landed before its barrier is armed again
      if (q >= LRF_ST_INFLIGHT) {
        const int j = q - LRF_ST_INFLIGHT;
        mbar_wait(a_full + 8 * (j & (kStR - 1)), (uint32_t)(j / kStR) & 1u, T.diag, 4u, (uint32_t)j, (uint32_t)q);
      }
      int u0 = 0, v0 = 0, u1, v1;
      bool real = p < gs;
      if (real) {
        box_of(k, u0, v0, u1, v1);
        // a box that lies entirely outside the lattice holds no node any lane will read: no TMA for it (the plane just counts as landed)
        real = u0 < gs && v0 < gs && u0 + W > 0 && v0 + H > 0;
      }
      if (real) {
        if (lane == 0) {
          sts_s32(a_base + 4 * slot, (int)dst - (v0 * W + u0) * kStRecB);
#if LRF_ST_DEBUG
          sts_s32(sbase + 448 + 8 * slot, k | (A << 16) | (q << 20));
          sts_s32(sbase + 452 + 8 * slot, (u0 & 0xffff) | (v0 << 16));
#endif
          mbar_expect_tx(bar, T.plane_bytes[A]);
          const int c1 = A == 0 ? k : u0, c2 = A == 0 ? u0 : (A == 1 ? k : v0), c3 = A == 2 ? k : v0;
          tma_load_4d(dst, &T.tm[A], bar, 0, c1, c2, c3);
        }
      } else if (lane == 0) {
#if LRF_ST_DEBUG
        sts_s32(sbase + 448 + 8 * slot, k | (A << 16) | (q << 20) | (1 << 19));
        sts_s32(sbase + 452 + 8 * slot, (u0 & 0xffff) | (v0 << 16));
#endif
        mbar_arrive(bar);
      }
      __syncwarp();

    }
    // no TMA may be in flight into this CTA's shared memory when the CTA retires: wait for the last min(q, R) planes issued
    for (int j = max(0, q - kStR); j < q; j++) mbar_wait(a_full + 8 * (j & (kStR - 1)), (uint32_t)(j / kStR) & 1u, T.diag, 2u, (uint32_t)j, (uint32_t)q);
    return;
  }

  // =========================== consumer warps ===========================
  uint32_t px, py;
  {
    const uint32_t q4 = lane >> 3, sub = lane & 7;
    px = tile_x0 + (warp & 1) * 8 + (q4 & 1) * 4 + (sub & 3);
    py = tile_y0 + (warp >> 1) * 4 + (q4 >> 1) * 2 + (sub >> 2);
  }
  const bool valid = px < w && py < h;
  float og[3], dg[3];
  gen_ray(S.view_inv, S.proj_inv, valid ? px : 0u, valid ? py : 0u, w, h, og, dg);
  bool alive = valid && fast_path_ok(dg) && ((pick3(dg, A) > 0.0f) == a_pos);
  uint32_t to_exact = (valid && !alive) ? 2u : 0u;                // 1: ambiguous break, 2: anything else the exact kernel has to do
  bool hit = false;
  // per-ray state, axes permuted to (U, V, A)
  float o_[3], d_[3], idir_[3], pa[3], tf[3], dpa[3];
  float M = 0.0f;
  uint32_t odd = 0, negm = 0, cnt = 0, seed_off = 0;
  int qe = 0;
  {
    int cur[3] = {0, 0, 0};
    if (alive) {
      const int e = staged_entry(S, og, dg, cur);
      if (e == 2) { alive = false; to_exact = 2u; }
      else if (e == 0) alive = false;                             // empty hit list: pixel 0 (ref :645,660)
      else hit = true;
    }
    const int perm[3] = {U, V, A};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float ov = pick3(og, perm[a]), dv = pick3(dg, perm[a]);
      const int c = pick3i(cur, perm[a]);
      const bool pos = dv > 0.0f;
      o_[a] = ov; d_[a] = dv;
      idir_[a] = 1.0f / dv;                                       // ref example_tracer.cpp:20-22
      dpa[a] = pos ? S.inv_gs_f : -S.inv_gs_f;
      pa[a] = (float)(c + (pos ? 1 : 0)) * S.inv_gs_f;            // exit plane of the first cell (exact: multiples of 2^-k)
      tf[a] = idir_[a] * (pa[a] - ov);
      const float tn = idir_[a] * ((float)(c + (pos ? 0 : 1)) * S.inv_gs_f - ov);
      M = a == 0 ? tn : fmaxf(M, tn);                             // entry time of the first cell == max of RayBoxIntersection's three min(lo,hi)
      odd |= (uint32_t)(c & 1) << a;
      negm |= (pos ? 0u : 1u) << a;
      cnt |= (0x200u | (uint32_t)(pos ? n - 1 - c : c)) << (10 * a);   // cells left along the axis under a guard bit
    }
    const int cA = pick3i(cur, A);
    qe = a_pos ? cA : n - 1 - cA;                                 // sequence number of the voxel's entry plane, counted from the face the tile marches in through
    seed_off = (uint32_t)((pick3i(cur, V) * W + pick3i(cur, U)) * kStRecB);
  }
  {
    const int first = __reduce_min_sync(FULL, alive ? qe : kStDone);
    if (lane == 0) { sts_s32(a_first + 4 * warp, first); sts_s32(a_prog + 4 * warp, first == kStDone ? kStDone : 0); }
  }
  __syncthreads();
  const bool ovf = lds_s32(a_ovf) != 0;
  int p0 = kStDone;
#pragma unroll
  for (int i = 0; i < NW; i++) p0 = min(p0, lds_s32(a_first + 4 * i));
  if (ovf) {                                                      // the tile's footprint does not fit the staged box: every pixel to the exact kernel
    if (threadIdx.x == 0 && T.stats) atomicAdd(T.stats + 2, 1u);
    if (valid) { alive = false; to_exact = 2u; }
  }
  qe -= p0;

  // SH basis of the ray as packed pairs (ref sh_eval_2 :83-106)
  u64 Y01, Y23, Y45, Y67;
  float Y8;
  {
    float Y[9];
    sh_eval_2(dg, Y);
    Y01 = pk2(Y[0], Y[1]); Y23 = pk2(Y[2], Y[3]); Y45 = pk2(Y[4], Y[5]); Y67 = pk2(Y[6], Y[7]);
    Y8 = Y[8];
  }
  const int dU2 = (negm & 1u) ? -2 * kStRecB : 2 * kStRecB, dV2 = ((negm & 2u) ? -2 * kStRecB : 2 * kStRecB) * W;

  // window: slot s = parity bits (U, V, A) of the node coordinates
#if !LRF_ST_RELOAD
  float4 wv[8];
#endif
  float wq[8];
  uint32_t AD[8];
#pragma unroll
  for (int s = 0; s < 8; s++) {
#if !LRF_ST_RELOAD
    wv[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#endif
    wq[s] = 0.0f; AD[s] = 0u;
  }
  int EA0 = 0, EA1 = 0;                                           // ring base of the window's plane with even / odd A coordinate
  uint32_t stale = 0x0fu;
  int phase = 0;                                                  // 0: window empty, 1: members J loaded, 2: steady state
  float thr = 1.0f, tPrev = -1.0f, col0 = 0.0f, col1 = 0.0f, col2 = 0.0f;
  float band = 0.0f;
  const float band_d = 0.11f * T.err_sigma;                       // per unit of delta
  uint32_t n_shaded = 0;
  int qready = 0, qpub = 0;
  uint32_t trip = 0;

  for (;;) {
    // ---- ring bookkeeping (warp-uniform) ----
    // A warp observes the planes strictly in sequence (qready) and publishes min(slowest lane, qready): the producer re-arms the
    // slot of plane q - R for plane q only when every live warp has BOTH seen q - R complete and marched past it.  So a slot's
    // barrier is never armed twice in one phase, and the parity a warp waits for always names the slot's current or previous use
    // (a warp whose rays enter the lattice late just walks through the earlier planes in the wait below).
    const bool can = alive && (qe + 3 <= qready);
    const bool any_can = __any_sync(FULL, can);
    if (!any_can || (trip++ & 1u) == 0u) {
    const int qmin = __reduce_min_sync(FULL, alive ? qe : kStDone);
    if (qmin == kStDone) break;
    const int pr = min(qmin, qready);
    if (pr != qpub) {
      qpub = pr;
      __syncwarp();
      if (lane == 0) sts_release_s32(a_prog + 4 * warp, pr);
    }
#if LRF_ST_DEBUG
    {
      const uint32_t ba = __ballot_sync(FULL, alive), bc = __ballot_sync(FULL, can);
      if (lane == 0) {
        const uint32_t d = sbase + 192 + 32 * warp;
        sts_s32(d, qmin); sts_s32(d + 4, qready); sts_s32(d + 8, (int)ba); sts_s32(d + 12, (int)bc); sts_s32(d + 16, qpub);
        sts_s32(d + 20, lds_s32(d + 20) + 1);
      }
    }
#endif
    if (!any_can) {
      mbar_wait(a_full + 8 * (qready & (kStR - 1)), (uint32_t)(qready / kStR) & 1u, T.diag, 3u, (uint32_t)qready, (uint32_t)qmin);
      qready++;
      continue;
    }
    // look-ahead (result first used by the next trip): keep the ready pointer ahead of the fastest lane
    // (the vote makes the answer the same in every lane: a phase that completes while the warp's test is in flight may be seen by some lanes only)
    if (qready < qmin + kStR && mbar_test_warp(a_full + 8 * (qready & (kStR - 1)), (uint32_t)(qready / kStR) & 1u)) qready++;
    }
    if (can) {
      if (phase == 0) {
        // seed the slot addresses of the first cell: slot s holds corner c = s ^ odd
        const int bEn = lds_s32(a_base + 4 * (qe & (kStR - 1))), bEx = lds_s32(a_base + 4 * ((qe + 1) & (kStR - 1)));
        const int bLo = a_pos ? bEn : bEx, bUp = a_pos ? bEx : bEn;
        const bool oA = (odd & 4u) != 0;
        EA0 = oA ? bUp : bLo; EA1 = oA ? bLo : bUp;
#pragma unroll
        for (int s = 0; s < 8; s++) {
          const uint32_t c = (uint32_t)s ^ odd;
          AD[s] = (uint32_t)((c & 4u) ? bUp : bLo) + seed_off + ((c & 1u) ? (uint32_t)kStRecB : 0u) + ((c & 2u) ? (uint32_t)(kStRecB * W) : 0u);
        }
        stale = 0x0fu;
      }
      // ---- (1) gather: pair J = slots {J, 7-J}; its stale member is J when bit J of `stale` is set, else 7-J ----
      // Chunk 0 {density, R, G, B} goes straight into the stale member's registers: two predicated loads from the two members'
      // OWN address registers (with one shared address ptxas folds them into one load + 8 predicated moves).
#if LRF_ST_RELOAD
      uint32_t ADc[8];                                             // chunk-0 addresses of the current cell's corners (the advance below moves AD on)
#pragma unroll
      for (int s8 = 0; s8 < 8; s8++) ADc[s8] = rec_chunk0(AD[s8]);
#endif
      const uint32_t st = stale;
      const uint32_t t0[4] = {st & 1u, (st >> 1) & 1u, (st >> 2) & 1u, (st >> 3) & 1u};
      float4 c1[4], c2[4];
      float c3[4];
#pragma unroll
      for (int J = 0; J < 4; J++) {
#if !LRF_ST_RELOAD
        lds_slot2(rec_chunk0(AD[J]), rec_chunk0(AD[7 - J]), t0[J], wv[J], wv[7 - J]);
#endif
#if LRF_ST_RELOAD
        const uint32_t r = t0[J] ? ADc[J] : ADc[7 - J];
#else
        const uint32_t r = rec_chunk0(t0[J] ? AD[J] : AD[7 - J]);
#endif
        c1[J] = lds_f4(rec_chunk(r, 1));
        c2[J] = lds_f4(rec_chunk(r, 2));
        c3[J] = lds_f1(rec_chunk(r, 3));
      }
      const bool shade = phase != 0;
      float lf0 = 0.0f, lf1 = 0.0f, lf2 = 0.0f, delta = 0.0f;
      const uint32_t odd_cur = odd;
      if (shade) {
        // ---- (2) which face does the ray leave through?  fp32 filter, exact fp64 when close (as lrf_fast.cuh) ----
        const float m = fminf(fminf(tf[0], tf[1]), tf[2]);
        const float gate = fmaf(fabsf(m), 0x1p-20f, 0x1p-100f);
        const bool c0 = tf[0] - m <= gate, cc1 = tf[1] - m <= gate, cc2 = tf[2] - m <= gate;
        bool is0 = c0, is1 = !c0 && cc1;
        if ((c0 && cc1) || (c0 && cc2) || (cc1 && cc2)) {
          const int cross = staged_exact_cross(pa, o_, d_);
          if (cross & (cross - 1)) { alive = false; to_exact = 2u; }   // edge / corner grazing: 3 or 7 boxes at the same distance
          is0 = (cross & 1) != 0;
          is1 = !is0 && (cross & 2) != 0;
        }
        const bool isA = !is0 && !is1;
        // ---- (3) geometry of the sample in the current cell, ref :134-153 (bit-exact) ----
        {
          const float tmin = M, tmax = m;
          if (tPrev < 0.0f) tPrev = tmin;                          // :139
          const float t = (tmin + tmax) * 0.5f;                    // :142
          delta = t - tPrev;                                       // :144
          if (delta > S.step_gate_f) delta = staged_long_step(delta, S.step);   // :145-146
          tPrev = t;                                               // :190
          // lower plane of the cell: the exit plane for a ray going down, one cell before it for a ray going up (exact either way)
          lf0 = ((o_[0] + t * d_[0]) - (pa[0] + fminf(-dpa[0], 0.0f))) * S.gs_f;   // :148-150
          lf1 = ((o_[1] + t * d_[1]) - (pa[1] + fminf(-dpa[1], 0.0f))) * S.gs_f;
          lf2 = ((o_[2] + t * d_[2]) - (pa[2] + fminf(-dpa[2], 0.0f))) * S.gs_f;
        }
        // ---- (4) advance to the next cell ----
        {
          const float e = is0 ? tf[0] : (is1 ? tf[1] : tf[2]);
          M = fmaxf(M, e);                                         // entry time of the next cell
          const bool sel[3] = {is0, is1, isA};
#pragma unroll
          for (int a = 0; a < 3; a++) {
            const float pan = pa[a] + dpa[a];                      // exact: multiples of 2^-k
            const float tfn = idir_[a] * (pan - o_[a]);
            tf[a] = sel[a] ? tfn : tf[a];
            pa[a] = sel[a] ? pan : pa[a];
          }
          const uint32_t abit = is0 ? 1u : (is1 ? 2u : 4u);
          const bool pst = ((odd ^ negm) & abit) != 0;             // parity of the node plane that enters the window
          odd ^= abit;
          cnt -= is0 ? 1u : (is1 ? (1u << 10) : (1u << 20));
          if ((cnt & 0x20080200u) != 0x20080200u) alive = false;   // left the lattice
          int dl = is0 ? dU2 : dV2;
          if (isA) {
            qe++;
            const int nb = lds_s32(a_base + 4 * ((qe + 1) & (kStR - 1)));
            dl = nb - (pst ? EA1 : EA0);
            EA0 = pst ? EA0 : nb; EA1 = pst ? nb : EA1;
          }
          // slot s changes iff its parity bit along the step axis equals pst; with exactly one axis stepping that is
          //   s = 0: !pst   s = 7: pst   s = 1: (is0 == pst)   s = 6: the opposite   s = 2 / 5: (is1 == pst)   s = 4 / 3: (isA == pst)
          const bool x1 = is0 == pst, x2 = is1 == pst, x4 = isA == pst;
          {
            const uint32_t d0 = pst ? 0u : (uint32_t)dl, d1 = x1 ? (uint32_t)dl : 0u, d2 = x2 ? (uint32_t)dl : 0u, d4 = x4 ? (uint32_t)dl : 0u;
            AD[0] += d0; AD[7] += (uint32_t)dl - d0;
            AD[1] += d1; AD[6] += (uint32_t)dl - d1;
            AD[2] += d2; AD[5] += (uint32_t)dl - d2;
            AD[4] += d4; AD[3] += (uint32_t)dl - d4;
          }
          stale = (pst ? 0u : 1u) | (x1 ? 2u : 0u) | (x2 ? 4u : 0u) | (x4 ? 0u : 8u);
        }
      } else stale = 0x00u;                                        // next: the other member of every pair
      // ---- (5) contract the gathered nodes with the ray's SH basis into the stale slots ----
#pragma unroll
      for (int J = 0; J < 4; J++) {
        u64 q2 = mul2(pk2(c1[J].x, c1[J].y), Y01);
        q2 = fma2(pk2(c1[J].z, c1[J].w), Y23, q2);
        q2 = fma2(pk2(c2[J].x, c2[J].y), Y45, q2);
        q2 = fma2(pk2(c2[J].z, c2[J].w), Y67, q2);
        float ql, qh;
        upk2(q2, ql, qh);
        const float qn = fmaf(c3[J], Y8, ql + qh);
        wq[J] = t0[J] ? qn : wq[J];
        wq[7 - J] = t0[J] ? wq[7 - J] : qn;
      }
      if (shade) {
        // ---- (6) interpolation weights per slot (the parity permutation applied to the weights) ----
        const bool oU = (odd_cur & 1u) != 0, oV = (odd_cur & 2u) != 0, oA = (odd_cur & 4u) != 0;
        const float gU = 1.0f - lf0, gV = 1.0f - lf1, gA = 1.0f - lf2;
        const float wu0 = oU ? lf0 : gU, wu1 = oU ? gU : lf0;
        const float wv0 = oV ? lf1 : gV, wv1 = oV ? gV : lf1;
        const float wa0 = oA ? lf2 : gA, wa1 = oA ? gA : lf2;
        const u64 pu0 = pk2(wu0, wu0), pu1 = pk2(wu1, wu1), pv0 = pk2(wv0, wv0), pv1 = pk2(wv1, wv1), pa0 = pk2(wa0, wa0), pa1 = pk2(wa1, wa1);
        const u64 w00 = mul2(pu0, pv0), w10 = mul2(pu1, pv0), w01 = mul2(pu0, pv1), w11 = mul2(pu1, pv1);
        const u64 w8[8] = {mul2(w00, pa0), mul2(w10, pa0), mul2(w01, pa0), mul2(w11, pa0), mul2(w00, pa1), mul2(w10, pa1), mul2(w01, pa1), mul2(w11, pa1)};
        // ---- (7) {density, R}, {G, B} as packed pairs, q scalar; two chains each ----
#if LRF_ST_RELOAD
        float4 wv[8];
#pragma unroll
        for (int s8 = 0; s8 < 8; s8++) wv[s8] = lds_f4(ADc[s8]);
#endif
        u64 dr0 = mul2(w8[0], pk2(wv[0].x, wv[0].y)), gb0 = mul2(w8[0], pk2(wv[0].z, wv[0].w));
        u64 dr1 = mul2(w8[1], pk2(wv[1].x, wv[1].y)), gb1 = mul2(w8[1], pk2(wv[1].z, wv[1].w));
        float ws[8], dump;
#pragma unroll
        for (int s = 0; s < 8; s++) upk2(w8[s], ws[s], dump);
        float q0 = ws[0] * wq[0], q1 = ws[1] * wq[1];
#pragma unroll
        for (int s = 2; s < 8; s += 2) {
          dr0 = fma2(w8[s], pk2(wv[s].x, wv[s].y), dr0); gb0 = fma2(w8[s], pk2(wv[s].z, wv[s].w), gb0);
          dr1 = fma2(w8[s + 1], pk2(wv[s + 1].x, wv[s + 1].y), dr1); gb1 = fma2(w8[s + 1], pk2(wv[s + 1].z, wv[s + 1].w), gb1);
          q0 = fmaf(ws[s], wq[s], q0); q1 = fmaf(ws[s + 1], wq[s + 1], q1);
        }
        float dn0, r0, g0, b0, dn1, r1, g1, b1;
        upk2(dr0, dn0, r0); upk2(gb0, g0, b0); upk2(dr1, dn1, r1); upk2(gb1, g1, b1);
        const float dens = dn0 + dn1, R = r0 + r1, G = g0 + g1, B = b0 + b1, Q = q0 + q1;
        const float sigma = dens < 0.0f ? 0.0f : dens;             // :170-171
        const float x = sigma * delta;
        const float tr = ex2_approx(x * -1.4426950408889634f);     // :175 (approximate; the break below knows by how much)
        const float wgt = thr * (1.0f - tr);                       // :187
        const float sc = lm_clamp(Q, 0.0f, 1.0f);                  // :181
        col0 = fmaf(wgt, lm_clamp(R * sc, 0.0f, 1.0f), col0);      // :182-187
        col1 = fmaf(wgt, lm_clamp(G * sc, 0.0f, 1.0f), col1);
        col2 = fmaf(wgt, lm_clamp(B * sc, 0.0f, 1.0f), col2);
        thr *= tr;                                                 // :189
        n_shaded++;
        // ---- the reference's break (:131) on the approximate throughput: decided only outside the error band ----
        band = fmaf(x, 0.11f * kStErrX, fmaf(fabsf(delta), band_d, band + 0.11f * kStErrStep));
        const float diff = thr - 0.1f;
        if (diff < -band) alive = false;                           // certainly below 0.1: the reference stops here too
        else if (!(diff >= band)) { alive = false; to_exact = to_exact ? to_exact : 1u; }   // too close to call (or NaN): the exact kernel renders this ray
      }
      phase = phase == 0 ? 1 : 2;
    }
  }
  __syncwarp();
  if (lane == 0) sts_release_s32(a_prog + 4 * warp, kStDone);

  if (T.stats) {
    const uint32_t b1 = __ballot_sync(FULL, to_exact == 1u), b2 = __ballot_sync(FULL, to_exact == 2u);
    if (lane == 0) {
      if (b1) atomicAdd(T.stats + 0, (uint32_t)__popc(b1));
      if (b2) atomicAdd(T.stats + 1, (uint32_t)__popc(b2));
    }
  }
  if (!valid) return;
  const size_t pixel = (size_t)py * w + px;
  if (to_exact) {
    fb_list[atomicAdd(fb_count, 1u)] = (uint32_t)pixel;
    if (tap_shaded) tap_shaded[pixel] = 0xffffffffu;
    return;
  }
  uint32_t pix = 0;                                                // ref :645,660: rays with an empty hit list pack float4(0)
  if (hit) pix = pack_color(lm_clamp(col0, 0.0f, 1.0f), lm_clamp(col1, 0.0f, 1.0f), lm_clamp(col2, 0.0f, 1.0f), 1.0f);   // :193,:653
  out[pixel] = pix;
  if (tap_shaded) tap_shaded[pixel] = n_shaded;
}

}  // namespace lrf
