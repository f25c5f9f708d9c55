// lrf_fast.cuh — the streamlined forward ray march (render + parity taps) for rays whose three direction
// components are all non-zero with a finite reciprocal (every ray of a perspective image except, at most, a
// centre row/column; the rest take the generic path of lrf_device.cuh through a fallback queue).
//
// Same results as the generic march (bit-exact hit set, order, shaded-sample count, transmittance; RGB to the
// 1e-4 tolerance), about a third of the instructions per sample.  What changed, and why it is still exact:
//
//  * Traversal decisions are taken on the fp32 slab times the shading needs anyway.  The reference shades
//    box (x,y,z) with  t_a(j) = (1.0f/d_a) * (j/gs - o_a)  for its two planes per axis (RayBoxIntersection,
//    ref example_tracer.cpp:18-43); it ORDERS boxes by the fp64 quotients  T_a(j) = (j/gs - o_a) / d_a
//    (IsRayHit, ref external/Octree/octree.h:449-488).  t_a(j) carries at most 3 roundings of 2^-24 relative
//    to T_a(j), so whenever the smallest exit time leads the other two by more than 2^-20 relative, the fp32
//    order IS the fp64 order and the ray crosses exactly one face: one new box.  Otherwise (about 1 step in
//    1000) the exact fp64 quotients are formed and compared — the reference's own comparison — and a true tie
//    (edge / corner grazing: 3 or 7 boxes, ordered by id) is expanded by build_event_cells like in the generic march.
//  * Per-axis slab times, plane coordinates and the 8 node offsets are carried incrementally: a step along axis a
//    touches only axis a's entry/exit time and the 4 nodes of the face that leaves.
//  * The 2x2x2 node window lives in registers, one slot per PARITY of the node coordinates (a face-adjacent
//    step replaces 4 nodes by 4 nodes of the same parities, so nothing moves between registers); stale slots
//    are tracked in an 8-bit mask and refilled by predicated loads.  The colour part keeps, per node,
//    {R, G, B, q = sum_j sh_j Y_j(d)} as two packed fp32 pairs and is contracted / interpolated with the
//    sm_100 packed-pair FMA (fma.rn.f32x2 -> FFMA2): this is the 1e-4-tolerance path, fused arithmetic is fine.
//    Everything that feeds a decision (density interpolation, expf, throughput product) stays op-for-op the
//    reference's unfused fp32 / fp64 sequence.
#pragma once
#include "lrf_device.cuh"

namespace lrf {

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// packed-pair fp32 (tolerance path only: ptxas is free to contract mul2 + add2 into FFMA2 even under -fmad=false)
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

struct FastOut
{
  float col[3], thr, cauchy;
  uint32_t n_hit, n_shaded;
};

// true when the ray can take the fast path
__device__ __forceinline__ bool fast_path_ok(const float *d)
{
  bool ok = true;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float inv = 1.0f / d[a];
    ok = ok && (d[a] != 0.0f) && (fabsf(inv) < __int_as_float(0x7f800000)) && (d[a] == d[a]);
  }
  return ok;
}

// MODE 0: render (stops at the reference's throughput break, ref example_tracer.cpp:131)
// MODE 1: parity taps (keeps traversing after the break to count every hit; records voxels, cauchy, touched nodes)
// MODE 2: backward, pass A — the forward again, keeping per-channel colour sums in fp64 and the shaded-sample count
// MODE 3: backward, pass B — the adjoint of every sample in forward order (closed form of the reverse sweep, see
//         lrf_kernels.cu), accumulated per window slot in shared memory and flushed with red.global.add when the slot's node
//         changes (gradient window: half the atomics of a per-sample scatter, 5 accumulators per node instead of 13)
#ifndef LRF_KBLOCK
#define LRF_KBLOCK 128            // threads per block of the thread-per-ray kernels.  Measured (r2r, variants33.log): sparse 512^3 render 463 / 451 / 432 Mrays/s
                                  // at 128 / 64 / 32 threads (same warps per SM), whole-image backward 12.0 ms at all three
#endif
constexpr int kFastBlock = LRF_KBLOCK;   // threads per block of the kernels that use FastMarch (shared-memory stride)
// Measured dead ends (profiles/r1f_variants.txt; the code is in the history of this file): taking the just-gathered nodes from
// registers instead of re-reading the shared window, the SH basis in registers, a rolling one-node-at-a-time gather, 96 or 168
// registers per thread, a branch-free single-basic-block iteration for dense grids (ptxas then schedules the gather
// consumers ahead of the traversal step and the latency cover is lost), and an L2 prefetch of the ray's coming nodes 2 / 4 / 8
// cells ahead (-2 %: the first-touch DRAM misses are not what the loop waits for).  Two more, both telling: replacing the exact
// density lerps + fp64 expf by 11 approximate instructions made the loop 7 % SLOWER (that arithmetic is what covers the second
// gather), and a rotated schedule that issues the next cell's pairs 0,1 before this cell's shading (twice the cover on paper)
// ran 36 % slower with the load-scoreboard stalls tripled (profiles/r1f_variants.txt boxes 8, 9); landing the next cell's nodes
// in shared memory through per-thread cp.async one sample ahead (no registers in flight) ran 2.2-2.8x slower (boxes 17, 18).
// What paid is the two-phase gather below.
#ifndef LRF_GRAD_MERGE
#define LRF_GRAD_MERGE 1     // image backward: lanes flushing the same node merge their adjoints before the reductions (0 = off; 1, 2, 4 = largest butterfly step)
#endif
#ifndef LRF_SEARCH_TRIPS
#define LRF_SEARCH_TRIPS 1   // sparse render: traversal events per lane between two shading slots of the warp (see run()).  4 was the optimum before the
                             // brick-distance jump shortened the searches; with it (and 3 blocks per SM) 1 / 2 / 3 / 4 events: 463 / 425 / 417 / 409 Mrays/s
#endif

// BAKED: the render reads the 64-byte node records {density, R, G, B | sh0..sh3 | sh4..sh7 | sh8, 0, 0, 0} (S.rec, lrf_staged.cuh) instead of the
// density + colour planes: a node is half a 128-byte line (the 48-byte colour records straddle lines) and its density comes with
// the first chunk, so a gather costs fewer L1 wavefronts (the pipe this kernel is bound by)
template <int MODE, bool SPARSE, bool GREY, bool BAKED = false> struct FastMarch
{
  const SceneDev &S;
  const uint64_t *etab;
  // ray
  float o_[3], d_[3], idir_[3];
  float4 *csm;       // this thread's colour window in shared memory: csm[slot * kFastBlock] = {R, G, B, q = sum_j sh_j Y_j(d)}
  float *ysm;        // this thread's SH basis in shared memory: ysm[j * kFastBlock], j = 0..8 (frees 9 registers)
  // traversal state of the CURRENT cell
  int cell[3];
  int sgn[3];
  float pa[3];       // coordinate of the plane through which the ray leaves the current cell
  float tn[3], tf[3];  // fp32 slab times of the entry / exit plane  ==  min(lo,hi) / max(lo,hi) of RayBoxIntersection
  // register window: slot s = (x parity) | (y parity) << 1 | (z parity) << 2 of the node coordinates
  uint32_t nx[2], ny[2], nz[2];   // node index contribution held by parity p of each axis: x, y*gs, z*gs^2
  float dn[8];
  uint32_t dirty;    // slots whose registers are stale
  // sparse sets: the coarse occupancy bit of the 8^3-voxel brick the ray is in, re-read only when the brick changes — in
  // empty space a traversal step then costs no memory access at all
  uint32_t brick_id = 0xffffffffu;   // index of the cached coarse word
  uint32_t brick_word = 0;
  uint32_t brick_b = 0xffffffffu;    // linear index of the brick whose distance byte is cached
  uint32_t far_dist = 0;             // Chebyshev distance (bricks) from the current cell's brick to the nearest occupied one; >= 2: the ray may jump (try_jump)
  bool far_empty = false;
  uint32_t odd;      // parity bits of the cell coordinates
  // compositing
  float thr = 1.0f, tPrev = -1.0f, cauchy = 0.0f;
  float col[3] = {0.0f, 0.0f, 0.0f};
  uint32_t n_hit = 0, n_shaded = 0;
  bool alive = true;
  // taps
  int32_t *vox = nullptr;
  uint32_t keep = 0;
  uint8_t *touched = nullptr;
  // backward (MODE 2 out / MODE 3 in)
  double csum[3] = {0.0, 0.0, 0.0};     // MODE 2: sum_k T_k (1 - tr_k) c_k per channel;  MODE 3: the running prefix of the same
  double ctot[3];                       // MODE 3: pass A's totals
  float aC[3], aTn, Tn;                 // MODE 3: adjoint of the clamped colour and of the final transmittance; the final transmittance
  uint32_t n_total = 0, k_done = 0;     // MODE 3: samples pass A shaded / samples processed so far
  float4 *gsm = nullptr;                // MODE 3: gradient window in shared memory, per slot {a_R, a_G, a_B, a_q} (stride kFastBlock)
  float *gdm = nullptr;                 //         ... density adjoint per slot
  uint32_t *nsm = nullptr;              //         ... node index the slot accumulates for

  __device__ FastMarch(const SceneDev &s, const uint64_t *t, float *ysm_, float4 *csm_) : S(s), etab(t), csm(csm_), ysm(ysm_) {}
  __device__ __forceinline__ float o(int a) const { return o_[a]; }
  __device__ __forceinline__ float d(int a) const { return d_[a]; }
  __device__ __forceinline__ float idir(int a) const { return idir_[a]; }

  __device__ __forceinline__ void init_ray(const float *of, const float *df)
  {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      o_[a] = of[a];
      d_[a] = df[a];
      idir_[a] = 1.0f / df[a];                      // ref example_tracer.cpp:20-22
      sgn[a] = df[a] > 0.0f ? 1 : -1;
    }
    if (!GREY) {
      float Y[9];
      sh_eval_2(df, Y);
#pragma unroll
      for (int j = 0; j < 9; j++) ysm[j * kFastBlock] = Y[j];
    }
  }

  __device__ __forceinline__ bool present_here()
  {
    if (S.occ8 == nullptr) return cell_present(S, cell[0], cell[1], cell[2]);   // dense set seen through a SPARSE instantiation (parity taps)
    // no branches: the coarse word is re-read only when its index changes, the fine word only inside an occupied brick —
    // both as PREDICATED loads, so that the lanes of a warp stay in lock-step through empty space
    const uint32_t b = (uint32_t)(cell[0] >> 3) + (uint32_t)(cell[1] >> 3) * S.nb8 + (uint32_t)(cell[2] >> 3) * S.nb8 * S.nb8;
    const uint32_t wi = b >> 5;
    uint32_t w = brick_word, fd = far_dist;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, %3;\n\t@p ld.global.nc.u32 %0, [%4];\n\tsetp.ne.u32 p, %5, %6;\n\t@p ld.global.nc.u8 %1, [%7];\n\t}"
                 : "+r"(w), "+r"(fd)
                 : "r"(wi), "r"(brick_id), "l"(S.occ8 + wi), "r"(b), "r"(brick_b), "l"(reinterpret_cast<const uint8_t *>(S.occ8 + S.occ8_words) + b));
    brick_word = w;
    brick_id = wi;
    brick_b = b;
    far_dist = fd;
    far_empty = fd >= 2u;
    const uint32_t on = (w >> (b & 31)) & 1u;
    const uint32_t lin = (uint32_t)cell[0] + (uint32_t)cell[1] * S.n + (uint32_t)cell[2] * S.n * S.n;
    uint32_t fw = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p ld.global.nc.u32 %0, [%2];\n\t}" : "+r"(fw) : "r"(on), "l"(S.occ + (lin >> 5)));
    return (fw >> (lin & 31)) & 1u;
  }
  __device__ __forceinline__ uint32_t slot_node(int s) const { return nx[s & 1] + ny[(s >> 1) & 1] + nz[s >> 2]; }

  // bring every piece of per-cell state to cell (x,y,z) from scratch (first cell, tie events, resync after them)
  __device__ __forceinline__ void set_cell(int x, int y, int z)
  {
    const int v[3] = {x, y, z};
    uint32_t nd[3][2];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      cell[a] = v[a];
      const int ex = v[a] + (sgn[a] > 0 ? 1 : 0), en = v[a] + (sgn[a] > 0 ? 0 : 1);
      pa[a] = (float)ex * S.inv_gs_f;
      tf[a] = idir(a) * (pa[a] - o(a));
      tn[a] = idir(a) * ((float)en * S.inv_gs_f - o(a));
      nd[a][0] = (uint32_t)(v[a] + (v[a] & 1));
      nd[a][1] = (uint32_t)(v[a] + 1 - (v[a] & 1));
    }
    const uint32_t gs = S.gs, gs2 = (uint32_t)S.gs * S.gs;
    nx[0] = nd[0][0]; nx[1] = nd[0][1];
    ny[0] = nd[1][0] * gs; ny[1] = nd[1][1] * gs;
    nz[0] = nd[2][0] * gs2; nz[1] = nd[2][1] * gs2;
    dirty = 0xffu;
    odd = (uint32_t)(x & 1) | ((uint32_t)(y & 1) << 1) | ((uint32_t)(z & 1) << 2);
  }

  // one face crossing along axis ax (select form: branch-free, the three axes' candidates are computed and the
  // crossed one is kept); returns the new cell index along that axis
  __device__ __forceinline__ int advance(int ax)
  {
    const bool m0 = ax == 0, m1 = ax == 1, m2 = ax == 2;
    const bool mm[3] = {m0, m1, m2};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float pan = pa[a] + (sgn[a] > 0 ? S.inv_gs_f : -S.inv_gs_f);   // exact: multiples of 2^-k
      const float tfn = idir(a) * (pan - o(a));
      tn[a] = mm[a] ? tf[a] : tn[a];
      tf[a] = mm[a] ? tfn : tf[a];
      pa[a] = mm[a] ? pan : pa[a];
      cell[a] = mm[a] ? cell[a] + sgn[a] : cell[a];
    }
    odd ^= (m0 ? 1u : 0u) | (m1 ? 2u : 0u) | (m2 ? 4u : 0u);
    const int csel = m0 ? cell[0] : (m1 ? cell[1] : cell[2]);
    const int ssel = m0 ? sgn[0] : (m1 ? sgn[1] : sgn[2]);
    // the node that enters lies on the new exit plane (coordinate csel + [ssel > 0]); it takes the slot of its parity
    const bool p1 = ((csel + (ssel > 0 ? 1 : 0)) & 1) != 0;
    const int stride2 = m0 ? 2 : (m1 ? 2 * S.gs : 2 * S.gs * S.gs);
    const uint32_t delta = (uint32_t)(ssel * stride2);
    nx[0] += (m0 && !p1) ? delta : 0u; nx[1] += (m0 && p1) ? delta : 0u;
    ny[0] += (m1 && !p1) ? delta : 0u; ny[1] += (m1 && p1) ? delta : 0u;
    nz[0] += (m2 && !p1) ? delta : 0u; nz[1] += (m2 && p1) ? delta : 0u;
    const uint32_t f0 = m0 ? 0x55u : (m1 ? 0x33u : 0x0fu);
    const uint32_t fm = p1 ? (f0 ^ 0xffu) : f0;
    dirty |= (m0 || m1 || m2) ? fm : 0u;
    return csel;
  }

  // Refill of the stale slots.  A face-adjacent step leaves exactly 4 stale slots and they never include both members
  // of an antipodal pair (s, 7 - s): a cube face contains one vertex of each of the 4 space diagonals.  So a refill
  // round runs over the 4 pairs: per pair ONE unconditional node load + SH contraction (no predicated-off work, no
  // divergent branches), written back into whichever member is stale.  After a full reset (both members stale) a
  // second round picks up the other members.
  // The round is split in two halves so that independent work sits between them: issue() starts the loads of the
  // cell about to be shaded, the march then does that cell's geometry and the whole traversal step to the NEXT cell
  // (~90 instructions that need no memory), and only then commit() consumes the loads.  That hides most of the
  // L2 latency of the gathers — the first thing the ncu stall breakdown of the unpipelined version pointed at —
  // at no cost in registers beyond the in-flight node data.
  struct NodeLoad
  {
    float td;
    float4 a, b, e;   // {R,G,B,sh0}, {sh1..sh4}, {sh5..sh8};  BAKED: {density,R,G,B}, {sh0..sh3}, {sh4..sh7} and s8
    float s8;
    uint32_t node;    // MODE 3: the node these values belong to (the gradient window needs it when the slot is flushed)
  };
  // bit J of the result: member 7-J of pair J is stale
  static __device__ __forceinline__ uint32_t high_members(uint32_t stale) { return __brev(stale) >> 24; }

  template <int J> __device__ __forceinline__ uint32_t pair_node(uint32_t stale) const
  {
    const bool wA = (stale >> J) & 1u;
    return wA ? slot_node(J) : slot_node(7 - J);
  }
  __device__ __forceinline__ void issue_colour(uint32_t node, NodeLoad &L) const
  {
    if (!GREY) {
      if (BAKED) {
        const float4 *pc = S.rec + (size_t)node * 4;
        L.b = __ldg(pc + 1); L.e = __ldg(pc + 2); L.s8 = __ldg(reinterpret_cast<const float *>(pc + 3));
      } else {
        const float4 *pc = S.col + (size_t)node * kColF4;
        L.a = __ldg(pc); L.b = __ldg(pc + 1); L.e = __ldg(pc + 2);
      }
    }
  }
  // the part of a node the density interpolation needs (BAKED: chunk 0, which carries R, G, B as well)
  __device__ __forceinline__ void issue_density(uint32_t node, NodeLoad &L) const
  {
    if (BAKED) L.a = __ldg(S.rec + (size_t)node * 4);
    else L.td = __ldg(S.dens + node);
  }
  template <int J> __device__ __forceinline__ void issue_pair(uint32_t stale, NodeLoad &L) const
  {
    const uint32_t node = pair_node<J>(stale);
    issue_density(node, L);
    if (MODE == 3) L.node = node;
    issue_colour(node, L);
  }
  // ---- gradient window (MODE 3) ----
  // Shared memory or registers, 2 / 3 / 4 blocks per SM: all 13.1-13.4 ms per 800x800 view at 256^3, and 7.0 ms with the red
  // instructions removed — the L2 atomic throughput is the bound (profiles/r1f_variants.txt boxes 10, 11).
  __device__ __forceinline__ void grad_window_reset()
  {
#pragma unroll
    for (int s = 0; s < 8; s++) {
      gsm[s * kFastBlock] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      gdm[s * kFastBlock] = 0.0f;
      nsm[s * kFastBlock] = 0u;
    }
  }
  // send slot s's accumulated adjoints to the gradient planes and clear them (what Enzyme's `grid_d +=` does per sample)
  __device__ __forceinline__ void grad_flush_slot(int s)
  {
    const uint32_t node = nsm[s * kFastBlock];
#if LRF_GRAD_MERGE
    // Neighbouring rays of a warp (pixels 0.2-0.5 voxel apart) walk the same voxel columns and reach this point with the SAME
    // node: the lanes that flush the same node here add their 13 adjoints with shuffles and one of them issues the four
    // reductions (the backward is bound by the L2 atomic throughput: 7.0 of its 12.6 ms remain when the red instructions are
    // removed).  Measured (profiles/r1f_variants.txt boxes 19, 20): partner lane ^ 1 only 13.1 -> 12.2 ms; adding lane ^ 2 /
    // lane ^ 4 costs more shuffles than it saves (12.4 / 13.0 ms); a __match_any_sync grouping with a serial shuffle loop over
    // the peers 22 ms (the groups are large, and merging 13 values per lane is ~150 instructions per node).
    float v[13];
    v[0] = gdm[s * kFastBlock];
    float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (!GREY) a = gsm[s * kFastBlock];
    v[1] = a.x; v[2] = a.y; v[3] = a.z;
#pragma unroll
    for (int j = 0; j < 9; j++) v[4 + j] = GREY ? 0.0f : ysm[j * kFastBlock] * a.w;
    gdm[s * kFastBlock] = 0.0f;
    if (!GREY) gsm[s * kFastBlock] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    // butterfly over the x-neighbours of the warp's 8x4 pixel tile (lane ^ 1, lane ^ 2[, lane ^ 4]): when the partner flushes the
    // same node the lower lane takes the sum and the upper one drops out
    const uint32_t active = __activemask();
    const int me = threadIdx.x & 31;
    uint32_t mynode = node;
#pragma unroll
    for (int step = 1; step <= LRF_GRAD_MERGE; step <<= 1) {
      const uint32_t pn = __shfl_xor_sync(active, mynode, step);
      const bool partner_on = (active >> (me ^ step)) & 1u;
      const bool same = partner_on && pn == mynode && mynode != 0xffffffffu;
      const bool upper = (me & step) != 0;
#pragma unroll
      for (int j = 0; j < (GREY ? 1 : 13); j++) {
        const float t = __shfl_xor_sync(active, v[j], step);
        if (same && !upper) v[j] += t;
      }
      if (same && upper) mynode = 0xffffffffu;                 // handed over
    }
    if (mynode == 0xffffffffu) return;
    if (v[0] != 0.0f) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(S.gdens + node), "f"(v[0]) : "memory");
    if (GREY) return;
    bool any = false;
#pragma unroll
    for (int j = 1; j < 13; j++) any = any || (v[j] != 0.0f);
    if (any) {
      float4 *q = S.gcol + (size_t)node * kColF4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + 1), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + 2), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]) : "memory");
    }
#else
    const float ad = gdm[s * kFastBlock];
    if (ad != 0.0f) {
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(S.gdens + node), "f"(ad) : "memory");
      gdm[s * kFastBlock] = 0.0f;
    }
    if (GREY) return;
    const float4 a = gsm[s * kFastBlock];
    if (a.x != 0.0f || a.y != 0.0f || a.z != 0.0f || a.w != 0.0f) {
      float4 *q = S.gcol + (size_t)node * kColF4;
      const float v = a.w;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(a.x), "f"(a.y), "f"(a.z), "f"(ysm[0] * v) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + 1), "f"(ysm[1 * kFastBlock] * v), "f"(ysm[2 * kFastBlock] * v),
                   "f"(ysm[3 * kFastBlock] * v), "f"(ysm[4 * kFastBlock] * v) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + 2), "f"(ysm[5 * kFastBlock] * v), "f"(ysm[6 * kFastBlock] * v),
                   "f"(ysm[7 * kFastBlock] * v), "f"(ysm[8 * kFastBlock] * v) : "memory");
      gsm[s * kFastBlock] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
#endif
  }
  __device__ __forceinline__ void grad_flush_all()
  {
#pragma unroll
    for (int s = 0; s < 8; s++) grad_flush_slot(s);
  }
  struct Basis
  {
    float Y0;                 // BAKED: Y8
    u64 Y12, Y34, Y56, Y78;   // BAKED: Y01, Y23, Y45, Y67
  };
  __device__ __forceinline__ void load_basis(Basis &Y) const
  {
    if (BAKED) {
      Y.Y0 = ysm[8 * kFastBlock];
      Y.Y12 = pk2(ysm[0], ysm[1 * kFastBlock]); Y.Y34 = pk2(ysm[2 * kFastBlock], ysm[3 * kFastBlock]);
      Y.Y56 = pk2(ysm[4 * kFastBlock], ysm[5 * kFastBlock]); Y.Y78 = pk2(ysm[6 * kFastBlock], ysm[7 * kFastBlock]);
      return;
    }
    Y.Y0 = ysm[0];
    Y.Y12 = pk2(ysm[1 * kFastBlock], ysm[2 * kFastBlock]); Y.Y34 = pk2(ysm[3 * kFastBlock], ysm[4 * kFastBlock]);
    Y.Y56 = pk2(ysm[5 * kFastBlock], ysm[6 * kFastBlock]); Y.Y78 = pk2(ysm[7 * kFastBlock], ysm[8 * kFastBlock]);
  }
  // density goes to its register slot; the colour, contracted with this ray's SH basis, to the slot's shared-memory
  // cell — the slot index is an address there, so no selects and no predicates are needed (a clean pair just rewrites
  // member 7-J with its own data)
  template <int J> __device__ __forceinline__ void grad_window_turn(uint32_t stale, uint32_t hi, uint32_t node)
  {
    const bool wA = (stale >> J) & 1u, wB = ((hi & ~stale) >> J) & 1u;
    if (wA || wB) {                                            // the slot's node really changes (a clean pair rewrites itself)
      const int slot = wA ? J : 7 - J;
      grad_flush_slot(slot);
      nsm[slot * kFastBlock] = node;
    }
  }
  template <int J> __device__ __forceinline__ void commit_density(uint32_t stale, uint32_t hi, float td)
  {
    const bool wA = (stale >> J) & 1u, wB = ((hi & ~stale) >> J) & 1u;
    dn[J] = wA ? td : dn[J];
    dn[7 - J] = wB ? td : dn[7 - J];
  }
  template <int J> __device__ __forceinline__ void commit_pair(uint32_t stale, uint32_t hi, const NodeLoad &L, const Basis &Y)
  {
    if (MODE == 3) grad_window_turn<J>(stale, hi, L.node);
    commit_density<J>(stale, hi, BAKED ? L.a.x : L.td);
    commit_colour<J>(stale, L, Y);
  }
  template <int J> __device__ __forceinline__ void commit_colour(uint32_t stale, const NodeLoad &L, const Basis &Y)
  {
    const bool wA = (stale >> J) & 1u;
    if (GREY) return;
    u64 q2 = mul2(pk2(L.b.x, L.b.y), Y.Y12);
    q2 = fma2(pk2(L.b.z, L.b.w), Y.Y34, q2);
    q2 = fma2(pk2(L.e.x, L.e.y), Y.Y56, q2);
    q2 = fma2(pk2(L.e.z, L.e.w), Y.Y78, q2);
    float ql, qh;
    upk2(q2, ql, qh);
    if (BAKED) {   // b, e hold sh0..sh7 (pairs with Y01..Y67 above), s8 the last coefficient; R, G, B ride in chunk 0 behind the density
      csm[(wA ? J : 7 - J) * kFastBlock] = make_float4(L.a.y, L.a.z, L.a.w, fmaf(L.s8, Y.Y0, ql + qh));
      return;
    }
    const float q = fmaf(L.a.w, Y.Y0, ql + qh);
    csm[(wA ? J : 7 - J) * kFastBlock] = make_float4(L.a.x, L.a.y, L.a.z, q);
  }
  __device__ __forceinline__ void refill_blocking(uint32_t stale)
  {
    while (stale) {
      const uint32_t hi = high_members(stale);
      NodeLoad L0, L1, L2, L3;
      issue_pair<0>(stale, L0); issue_pair<1>(stale, L1); issue_pair<2>(stale, L2); issue_pair<3>(stale, L3);
      Basis Y;
      if (!GREY) load_basis(Y);
      commit_pair<0>(stale, hi, L0, Y); commit_pair<1>(stale, hi, L1, Y); commit_pair<2>(stale, hi, L2, Y); commit_pair<3>(stale, hi, L3, Y);
      // served this round: the low member of each pair if stale, else the high member
      const uint32_t lowm = stale & 0x0fu;
      stale &= ~(lowm | (__brev((hi & ~lowm & 0x0fu)) >> 24));
    }
  }

  // geometry of the sample in the current cell, ref RayGridIntersection example_tracer.cpp:134-153 (bit-exact)
  struct Geom
  {
    float delta, lf[3];
    uint32_t odd;   // parity bits of the cell coordinates
  };
  __device__ __forceinline__ void geometry(Geom &g)
  {
    const float tmin = fmaxf(fmaxf(tn[0], tn[1]), tn[2]);      // all finite: == the std::max / std::min chains of :18-43
    const float tmax = fminf(fminf(tf[0], tf[1]), tf[2]);
    if (tPrev < 0.0f) tPrev = tmin;                            // :139
    const float t = (tmin + tmax) * 0.5f;                      // :142
    float delta = t - tPrev;                                   // :144
    if (delta > S.step_gate_f)                                 // :145-146
      delta = (float)((double)delta - (double)((int)((double)delta / S.step) - 1) * S.step);
    tPrev = t;                                                 // :190
    g.delta = delta;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float bmin = sgn[a] > 0 ? pa[a] - S.inv_gs_f : pa[a];   // == (float)cell / (float)gs, exact
      const float p = o(a) + t * d(a);                         // :148
      g.lf[a] = (p - bmin) * S.gs_f;                           // :150
    }
    g.odd = odd;
  }

  // shade one sample from the window; false = the reference's throughput break fired
  struct Sig
  {
    float tr, wgt;
    float dens;   // interpolated density before the relu (the backward needs its sign and, as sigma, its value)
  };
  // density part of a sample (decision path, bit-exact): transmittance of the step and the compositing weight
  __device__ __forceinline__ Sig shade_sigma(const Geom &g)
  {
    const bool od[3] = {(g.odd & 1u) != 0, (g.odd & 2u) != 0, (g.odd & 4u) != 0};
    const float *lf = g.lf;
    // ---- density: nested lerps in the reference's order (:155-167 via lerpCell :72-80), bit-exact ----
    float A[4], B[2];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float p0 = dn[2 * j], p1 = dn[2 * j + 1];
      A[j] = lm_lerp(od[0] ? p1 : p0, od[0] ? p0 : p1, lf[0]);
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const float p0 = A[2 * j], p1 = A[2 * j + 1];
      B[j] = lm_lerp(od[1] ? p1 : p0, od[1] ? p0 : p1, lf[1]);
    }
    float sigma = lm_lerp(od[2] ? B[1] : B[0], od[2] ? B[0] : B[1], lf[2]);
    Sig r;
    r.dens = sigma;
    if (sigma < 0.0f) sigma = 0.0f;                            // :170-171
    if (MODE == 1) cauchy += logf(1.0f + 2.0f * sigma * sigma);  // :173
    r.tr = expf_libm(-sigma * g.delta, etab);                  // :175
    r.wgt = thr * (1.0f - r.tr);                               // :187
    return r;
  }
  // interpolation weight of every window slot (the parity permutation applied to the weights instead of the values)
  __device__ __forceinline__ void slot_weights(const Geom &g, float *w) const
  {
    const bool od[3] = {(g.odd & 1u) != 0, (g.odd & 2u) != 0, (g.odd & 4u) != 0};
    const float *lf = g.lf;
    const float gx = 1.0f - lf[0], gy = 1.0f - lf[1], gz = 1.0f - lf[2];
    const float wx0 = od[0] ? lf[0] : gx, wx1 = od[0] ? gx : lf[0];
    const float wy0 = od[1] ? lf[1] : gy, wy1 = od[1] ? gy : lf[1];
    const float wz0 = od[2] ? lf[2] : gz, wz1 = od[2] ? gz : lf[2];
    const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
    w[0] = w00 * wz0; w[1] = w10 * wz0; w[2] = w01 * wz0; w[3] = w11 * wz0;
    w[4] = w00 * wz1; w[5] = w10 * wz1; w[6] = w01 * wz1; w[7] = w11 * wz1;
  }
  // MODE 3: the adjoint of one sample, ref RayGridLoss / RayGridIntersection differentiated by hand (the same
  // formulas as BwdVisitor in lrf_kernels.cu, which the oracle's orc_render_grad restates on the CPU).  false = last sample.
  __device__ __forceinline__ bool shade_backward(const Geom &g, const Sig sg)
  {
    const bool last = (k_done + 1 == n_total);
    k_done++;
    if (sg.dens < 0.0f) return !last;                          // relu closed: sigma = 0, tr = 1, weight 0 — no adjoint reaches the grid (:170-171)
    const float sigma = sg.dens, tr = sg.tr, wgt = sg.wgt;
    float w[8];
    slot_weights(g, w);
    // the colour is needed even when wgt == 0: d colour / d tr = -T_k c_k does not vanish with (1 - tr)
    float c[3] = {1.0f, 1.0f, 1.0f}, acc[3] = {0.0f, 0.0f, 0.0f}, raw = 0.0f, s = 0.0f;
    if (!GREY) {
      float R = 0.0f, G = 0.0f, Bc = 0.0f, Q = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float4 cv = csm[k * kFastBlock];
        R = fmaf(w[k], cv.x, R); G = fmaf(w[k], cv.y, G); Bc = fmaf(w[k], cv.z, Bc); Q = fmaf(w[k], cv.w, Q);
      }
      acc[0] = R; acc[1] = G; acc[2] = Bc;
      raw = Q;
      s = lm_clamp(Q, 0.0f, 1.0f);                             // :181
      c[0] = lm_clamp(R * s, 0.0f, 1.0f); c[1] = lm_clamp(G * s, 0.0f, 1.0f); c[2] = lm_clamp(Bc * s, 0.0f, 1.0f);
    }
    if (wgt != 0.0f) {
#pragma unroll
      for (int ch = 0; ch < 3; ch++) csum[ch] += (double)(wgt * c[ch]);   // prefix INCLUDING this sample
    }
    const float dot = c[0] * aC[0] + c[1] * aC[1] + c[2] * aC[2];
    float a_tr;
    if (last) a_tr = -thr * dot + aTn * thr;
    else {
      const double suffix = (ctot[0] - csum[0]) * (double)aC[0] + (ctot[1] - csum[1]) * (double)aC[1] + (ctot[2] - csum[2]) * (double)aC[2];
      a_tr = -thr * dot + (float)((suffix + (double)aTn * (double)Tn) / (double)tr);
    }
    // d tr / d sigma = -delta tr;  cauchy: d/dsigma log(1 + 2 sigma^2) = 4 sigma / (1 + 2 sigma^2), weight 0.0005 (:267)
    const float a_dens = -g.delta * tr * a_tr + 0.0005f * (4.0f * sigma / (1.0f + 2.0f * sigma * sigma));
    float a_rgb[3] = {0.0f, 0.0f, 0.0f}, a_raw = 0.0f;
    if (!GREY && wgt != 0.0f) {
      float a_s = 0.0f;
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const float a_c = wgt * aC[ch];
        const float u = acc[ch] * s;
        if (0.0f < u && u <= 1.0f) { a_rgb[ch] = s * a_c; a_s = fmaf(acc[ch], a_c, a_s); }   // clamp passes iff 0 < u <= 1
      }
      a_raw = (0.0f < raw && raw <= 1.0f) ? a_s : 0.0f;
    }
    if (a_dens != 0.0f) {
#pragma unroll
      for (int k = 0; k < 8; k++) gdm[k * kFastBlock] = fmaf(w[k], a_dens, gdm[k * kFastBlock]);
    }
    if (a_raw != 0.0f || a_rgb[0] != 0.0f || a_rgb[1] != 0.0f || a_rgb[2] != 0.0f) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        float4 a = gsm[k * kFastBlock];
        a.x = fmaf(w[k], a_rgb[0], a.x); a.y = fmaf(w[k], a_rgb[1], a.y); a.z = fmaf(w[k], a_rgb[2], a.z); a.w = fmaf(w[k], a_raw, a.w);
        gsm[k * kFastBlock] = a;
      }
    }
    thr *= tr;                                                 // :189
    return !last;
  }
  // colour part (1e-4-tolerance path) + the throughput update; false = the reference's throughput break fired
  __device__ __forceinline__ bool shade_colour(const Geom &g, const Sig sg)
  {
    if (MODE == 3) return shade_backward(g, sg);
    const bool od[3] = {(g.odd & 1u) != 0, (g.odd & 2u) != 0, (g.odd & 4u) != 0};
    const float *lf = g.lf;
    const float tr = sg.tr, wgt = sg.wgt;
    if (wgt != 0.0f) {                                         // (shading zero-weight samples too instead of branching pays in the register-window kernel, not here: 2.33 -> 2.47 ms)
      if (GREY) {
        col[0] = col[0] + wgt; col[1] = col[1] + wgt; col[2] = col[2] + wgt;
        if (MODE == 2) { csum[0] += (double)wgt; csum[1] += (double)wgt; csum[2] += (double)wgt; }
      } else {
        // ---- colour: 1e-4-tolerance path, packed-pair FMAs ----
        const float gx = 1.0f - lf[0], gy = 1.0f - lf[1], gz = 1.0f - lf[2];
        const float wx0 = od[0] ? lf[0] : gx, wx1 = od[0] ? gx : lf[0];
        const float wy0 = od[1] ? lf[1] : gy, wy1 = od[1] ? gy : lf[1];
        const float wz0 = od[2] ? lf[2] : gz, wz1 = od[2] ? gz : lf[2];
        const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
        const float w[8] = {w00 * wz0, w10 * wz0, w01 * wz0, w11 * wz0, w00 * wz1, w10 * wz1, w01 * wz1, w11 * wz1};
        // two independent accumulation chains per pair of channels (latency), merged at the end
        const float4 c0 = csm[0], c1 = csm[kFastBlock];
        u64 RG0 = mul2(pk2(w[0], w[0]), pk2(c0.x, c0.y)), BQ0 = mul2(pk2(w[0], w[0]), pk2(c0.z, c0.w));
        u64 RG1 = mul2(pk2(w[1], w[1]), pk2(c1.x, c1.y)), BQ1 = mul2(pk2(w[1], w[1]), pk2(c1.z, c1.w));
#pragma unroll
        for (int k = 2; k < 8; k += 2) {
          const float4 ca = csm[k * kFastBlock], cb = csm[(k + 1) * kFastBlock];
          const u64 wk = pk2(w[k], w[k]), wk1 = pk2(w[k + 1], w[k + 1]);
          RG0 = fma2(wk, pk2(ca.x, ca.y), RG0);
          BQ0 = fma2(wk, pk2(ca.z, ca.w), BQ0);
          RG1 = fma2(wk1, pk2(cb.x, cb.y), RG1);
          BQ1 = fma2(wk1, pk2(cb.z, cb.w), BQ1);
        }
        float R, G, Bc, Q, R1, G1, B1, Q1;
        upk2(RG0, R, G); upk2(BQ0, Bc, Q); upk2(RG1, R1, G1); upk2(BQ1, B1, Q1);
        R += R1; G += G1; Bc += B1; Q += Q1;
        const float s = lm_clamp(Q, 0.0f, 1.0f);               // :181
        if (MODE == 2) {                                       // backward pass A: the terms also go into the fp64 sums
          const float t0 = wgt * lm_clamp(R * s, 0.0f, 1.0f), t1 = wgt * lm_clamp(G * s, 0.0f, 1.0f), t2 = wgt * lm_clamp(Bc * s, 0.0f, 1.0f);
          col[0] = col[0] + t0; col[1] = col[1] + t1; col[2] = col[2] + t2;
          csum[0] += (double)t0; csum[1] += (double)t1; csum[2] += (double)t2;
        } else {
          col[0] = fmaf(wgt, lm_clamp(R * s, 0.0f, 1.0f), col[0]);  // :182-187
          col[1] = fmaf(wgt, lm_clamp(G * s, 0.0f, 1.0f), col[1]);
          col[2] = fmaf(wgt, lm_clamp(Bc * s, 0.0f, 1.0f), col[2]);
        }
      }
    }
    thr *= tr;                                                 // :189
    n_shaded++;
    return !(thr < 0.1f);                                      // :131
  }

  // Empty-space jump for sparse sets.  Every 8^3-voxel brick carries its Chebyshev distance k (in bricks) to the nearest occupied
  // brick.  k >= 2: the cube of bricks within k - 2 of the current one AND its one-brick rim are empty, so nothing the ray can touch
  // before it has left that cube — face crossings, or the 3 / 7 cells of an edge / corner event on the cube's boundary — is present.
  // So the traversal is re-seeded in the cell that holds the ray point just behind the cube's exit plane, PROVIDED that cell is
  // certainly on the exact path: its fp32 slab interval
  // [max tn, min tf] must be open by more than the rounding bound (then the exact fp64 interval is non-empty and the ray passes
  // through the cell's interior; the per-cell state is a function of the cell alone, so the march continues bit for bit as if it
  // had stepped there).  Otherwise nothing is changed and the march steps on voxel by voxel.
  // Returns -1: the ray leaves the lattice (nothing present is left on it), 0: no jump, 1: jumped.
  __device__ __forceinline__ int try_jump()
  {
    float tb = __int_as_float(0x7f800000);
    const int r8 = ((int)far_dist - 2) * 8;                    // the bricks within far_dist - 2 of this one form the box to leave; its rim is empty too
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int lo = cell[a] & ~7;
      const float pb = (float)(sgn[a] > 0 ? lo + 8 + r8 : lo - r8) * S.inv_gs_f;
      tb = fminf(tb, idir(a) * (pb - o(a)));
    }
    const float tj = tb + fmaf(fabsf(tb), 0x1p-17f, 0x1p-40f);
    int c[3];
    bool outside = false;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      c[a] = (int)floorf((o(a) + tj * d(a)) * S.gs_f);
      outside = outside || (unsigned)c[a] >= (unsigned)S.n;
    }
    if (outside) return -1;
    if (c[0] == cell[0] && c[1] == cell[1] && c[2] == cell[2]) return 0;
    const int old[3] = {cell[0], cell[1], cell[2]};
    const uint32_t old_dirty = dirty;
    set_cell(c[0], c[1], c[2]);
    const float tmin = fmaxf(fmaxf(tn[0], tn[1]), tn[2]), tmax = fminf(fminf(tf[0], tf[1]), tf[2]);
    if (tmax - tmin > fmaf(fabsf(tmax), 0x1p-19f, 0x1p-100f)) return 1;
    set_cell(old[0], old[1], old[2]);                          // too close to call: back to the cell we stood on (same state, bit for bit)
    dirty = old_dirty | dirty;
    return 0;
  }

  // fp32 filter on the three exit times: the axis whose plane comes first, or -1 when two of them are within the
  // rounding bound of each other (then the exact fp64 quotients decide)
  __device__ __forceinline__ int decide_fast() const
  {
    const float m = fminf(fminf(tf[0], tf[1]), tf[2]);
    const float gate = fmaf(fabsf(m), 0x1p-20f, 0x1p-100f);
    const bool c0 = tf[0] - m <= gate, c1 = tf[1] - m <= gate, c2 = tf[2] - m <= gate;
    const int ax = c0 ? 0 : (c1 ? 1 : 2);
    return ((c0 && c1) || (c0 && c2) || (c1 && c2)) ? -1 : ax;
  }

  // exact comparison of the three exit-plane times (the reference's own fp64 quotients); returns the crossing mask
  __device__ __forceinline__ int exact_cross() const
  {
    double te[3];
#pragma unroll
    for (int a = 0; a < 3; a++) te[a] = ((double)pa[a] - (double)o(a)) / (double)d(a);
    const double em = fmin(te[0], fmin(te[1], te[2]));
    return (te[0] == em ? 1 : 0) | (te[1] == em ? 2 : 0) | (te[2] == em ? 4 : 0);
  }

  __device__ void run(const float *of, const float *df, uint32_t ws_mask = 0xffffffffu)
  {
    init_ray(of, df);
    const int n = S.n;
    // ---- entry: exact fp64, once per ray (same construction as the generic march) ----
    uint32_t queue[8];
    int nq = 0;
    int cur[3];
    bool leaving_after = false;
    {
      const double INF = __longlong_as_double(0x7ff0000000000000LL);
      double t_enter = -INF, t_exit = INF, od[3], hd[3];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        od[a] = (double)of[a];
        hd[a] = (double)df[a];
        const double T0 = (0.0 - od[a]) / hd[a], Tn = ((double)n * S.inv_gs - od[a]) / hd[a];
        t_enter = fmax(t_enter, sgn[a] > 0 ? T0 : Tn);
        t_exit = fmin(t_exit, sgn[a] > 0 ? Tn : T0);
      }
      const double t_start = fmax(t_enter, 0.0);
      if (!(t_exit >= t_start)) return;                        // ref octree.h:483 for the whole lattice
      int lo[3], two_mask = 0;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const int s = sgn[a];
        auto Tex = [&](int j) { return ((double)j * S.inv_gs - od[a]) / hd[a]; };
        int i = (int)floor((od[a] + t_start * hd[a]) * (double)S.gs);
        i = max(0, min(n - 1, i));
        for (int it = 0; it < 4; it++) {
          if (Tex(i + (s > 0 ? 0 : 1)) > t_start && (unsigned)(i - s) < (unsigned)n) { i -= s; continue; }
          if (Tex(i + (s > 0 ? 1 : 0)) < t_start && (unsigned)(i + s) < (unsigned)n) { i += s; continue; }
          break;
        }
        cur[a] = i;
        lo[a] = i;
        if (Tex(i + (s > 0 ? 0 : 1)) == t_start && (unsigned)(i - s) < (unsigned)n) {
          two_mask |= 1 << a;
          lo[a] = min(i, i - s);
        } else if (Tex(i + (s > 0 ? 1 : 0)) == t_start && (unsigned)(i + s) < (unsigned)n) {
          two_mask |= 1 << a;
          lo[a] = min(i, i + s);
          cur[a] = i + s;
        }
      }
      if (two_mask == 0) {
        if (cell_present(S, lo[0], lo[1], lo[2])) queue[nq++] = pack_cell(lo[0], lo[1], lo[2]);
      } else {
        uint32_t cells[8];
        const int m = build_event_cells(S, lo[0], lo[1], lo[2], two_mask, 0, 0, cells);
        for (int j = 0; j < m; j++) queue[j] = cells[m - 1 - j];   // popped from the back
        nq = m;
      }
    }
    const int pos_mask = (sgn[0] > 0 ? 1 : 0) | (sgn[1] > 0 ? 2 : 0) | (sgn[2] > 0 ? 4 : 0);
    bool resync = true;
    cell[0] = cell[1] = cell[2] = -8;
    dirty = 0;

    // ONE traversal event: +1 = standing on a cell to be shaded, 0 = keep going (empty voxel of a sparse set, or a tie event
    // was queued), -1 = the ray is done
    auto step_once = [&]() -> int {
      if (nq > 0) {
        int x, y, z;
        unpack_cell(queue[--nq], x, y, z);
        set_cell(x, y, z);
        far_empty = false;                                     // describes another cell's brick
        return 1;
      }
      if (resync) {
        if (leaving_after) return -1;
        if (cell[0] != cur[0] || cell[1] != cur[1] || cell[2] != cur[2]) set_cell(cur[0], cur[1], cur[2]);
        resync = false;
        far_empty = false;
      }
      if (SPARSE && far_empty) {                               // empty space: to the exit of the brick in one go
        const int j = try_jump();
        if (j < 0) return -1;
        if (j > 0) return present_here() ? 1 : 0;
      }
      // ---- which face does the ray leave through?  fp32 filter, exact fp64 when close ----
      int ax = decide_fast();
      if (ax < 0) {
        const int cross = exact_cross();
        if (cross & (cross - 1)) {
          // true tie: the ray passes through a lattice edge / corner — 3 or 7 boxes at the same distance
          int lo[3];
          leaving_after = false;
#pragma unroll
          for (int a = 0; a < 3; a++) {
            lo[a] = cell[a];
            cur[a] = cell[a];
            if (cross & (1 << a)) {
              const int nw = cell[a] + sgn[a];
              lo[a] = min(cell[a], nw);
              cur[a] = nw;
              if ((unsigned)nw >= (unsigned)n) leaving_after = true;
            }
          }
          uint32_t cells[8];
          const int mm = build_event_cells(S, lo[0], lo[1], lo[2], cross, cross, pos_mask, cells);
          for (int j = 0; j < mm; j++) queue[j] = cells[mm - 1 - j];
          nq = mm;
          resync = true;
          return 0;
        }
        ax = (cross & 1) ? 0 : ((cross & 2) ? 1 : 2);
      }
      const int moved = advance(ax);
      if ((unsigned)moved >= (unsigned)n) return -1;           // left the lattice
      if (SPARSE && !present_here()) return 0;
      return 1;
    };
    // moves the traversal to the next cell to be shaded; false = the ray is done
    auto next_cell = [&]() -> bool {
      for (;;) {
        const int st = step_once();
        if (st) return st > 0;
      }
    };

    // One sample: (1) start the gathers of the current cell's stale nodes — pairs 0,1 complete, pairs 2,3 density only —,
    // (2) the cell's geometry and the whole traversal step to the next cell (no memory needed: this is what covers the
    // gathers), (3) commit pairs 0,1, start the colour of pairs 2,3 behind the density lerps + the fp64 expf chain, commit
    // them, colour, composite.  The two-phase gather keeps 26 fewer registers live across the traversal step: 60 -> 16 bytes
    // of spills at 128 registers, 2.70 -> 2.26 ms per 800x800 view (profiles/r1f_variants.txt).
    struct Pending
    {
      uint32_t stale, hi, node2, node3;
      NodeLoad L0, L1, L2, L3;
      Geom g;
    };
    auto sample_begin = [&](Pending &P) {
      P.stale = dirty;
      P.hi = high_members(P.stale);
      if (P.stale & P.hi & 0x0fu) {                            // both members of a pair stale (after a full reset): rare
        refill_blocking(P.stale);
        P.stale = 0;
        P.hi = 0;
      }
      P.node2 = pair_node<2>(P.stale); P.node3 = pair_node<3>(P.stale);
      issue_pair<0>(P.stale, P.L0); issue_pair<1>(P.stale, P.L1);
      issue_density(P.node2, P.L2); issue_density(P.node3, P.L3);
      dirty = 0;
      geometry(P.g);
    };
    auto sample_end = [&](Pending &P) -> bool {               // false = the reference's throughput break fired
      Basis Y;
      if (!GREY) load_basis(Y);
      commit_pair<0>(P.stale, P.hi, P.L0, Y); commit_pair<1>(P.stale, P.hi, P.L1, Y);
      if (MODE == 3) { grad_window_turn<2>(P.stale, P.hi, P.node2); grad_window_turn<3>(P.stale, P.hi, P.node3); }
      commit_density<2>(P.stale, P.hi, BAKED ? P.L2.a.x : P.L2.td); commit_density<3>(P.stale, P.hi, BAKED ? P.L3.a.x : P.L3.td);
      issue_colour(P.node2, P.L2); issue_colour(P.node3, P.L3);   // in flight across the density lerps + expf chain
      const Sig sg = shade_sigma(P.g);
      commit_colour<2>(P.stale, P.L2, Y); commit_colour<3>(P.stale, P.L3, Y);
      return shade_colour(P.g, sg);
    };

    if (SPARSE && MODE == 0) {
      // Sparse render.  A lane that is looking for its next occupied voxel must not hold up the lanes that stand on one
      // (a ray leaving the sphere of config 4 steps through ~200 empty voxels while its neighbours shade), and the search
      // itself must stay in lock-step.  So every lane gets LRF_SEARCH_TRIPS traversal events per trip of this loop, taken
      // warp-synchronously (the vote re-converges the warp after every event; without it the lanes that take a rare path —
      // an exact-compare step, a coarse-word reload — never re-merge before the search ends: ncu showed 7.8 active lanes per
      // instruction, profiles/r1k_sparse_ncu_summary.txt), and the lanes that stand on a voxel shade it in the same trip.
      // ws_mask = the lanes of this warp that march a ray in this call (a ballot taken by the kernel at a converged point).
      // A lane whose ray is finished returns and exits the kernel: exited lanes are not waited for by the votes.
      auto search = [&]() -> int {
        int st = 0, trips = LRF_SEARCH_TRIPS;
        do {
          if (st == 0) st = step_once();
        } while (--trips > 0 && __any_sync(ws_mask, st == 0));
        return st;
      };
      int st = search();
      for (;;) {
        if (st < 0) return;
        const bool have = st > 0;
        Pending P;
        if (have) {
          n_hit = 1;
          sample_begin(P);
        }
        st = search();                                         // on a voxel: step to the next one; searching: keep searching
        if (have && !sample_end(P)) return;
      }
    }

    if (!next_cell()) return;
    for (;;) {
      // ---- the current cell is a hit ----
      if (MODE == 1) {
        if (vox && n_hit < keep) { vox[n_hit * 3] = cell[0]; vox[n_hit * 3 + 1] = cell[1]; vox[n_hit * 3 + 2] = cell[2]; }
        n_hit++;
        if (!alive) {                                          // past the break: only count the remaining hits
          if (!next_cell()) return;
          continue;
        }
        if (touched) {
#pragma unroll
          for (int s = 0; s < 8; s++) touched[slot_node(s)] = 1;
        }
      } else n_hit = 1;
      Pending P;
      sample_begin(P);
      const bool more = next_cell();
      if (!sample_end(P)) {
        alive = false;
        if (MODE != 1) return;
      }
      if (!more) return;
    }
  }
};

}  // namespace lrf
