// lrf_lean.cuh — the dense colour forward march with the whole 2x2x2 node window in REGISTERS (round 2).
//
// Same results as FastMarch<0> of lrf_fast.cuh (bit-exact hit set, order, shaded-sample count and transmittance: every
// decision-critical operation is the same IEEE operation on the same operands; RGB on the 1e-4 tolerance path), 9 % fewer
// instructions, a quarter fewer L1 wavefronts and three fewer divergent branches per sample.  ncu on FastMarch
// (profiles/r1t_render_ncu_summary.txt) showed the L1 data pipe 88 % busy and 415 warp instructions per sample-iteration; 57 of its
// 154 wavefronts per iteration were the shared-memory colour window and the SH basis, 16 % of its instructions the selects of the
// parity-slot permutation and of a traversal step that computed all three axes.  What is different here:
//
//  * the colour window {R, G, B, sh0} x 8, q' = sum_{j>=1} sh_j Y_j(d) x 8 and the SH basis live in registers.  A face step replaces
//    one member of each antipodal slot pair (J, 7 - J); the member is chosen by PREDICATED loads that write straight into the slot's
//    registers (@p ld [A] -> slot J, @!p ld [B] -> slot 7 - J: two issue slots, no selects, no shared-memory round trip), and the
//    SH contraction's last add is predicated the same way.  Nothing reads a loaded register before the shading (a move of .w right
//    behind the load made the warp wait for the gather there and then: 3.2 ms instead of 2.2).
//  * the slab entry time is ONE running maximum: the reference's tmin = max_a tn[a] (RayBoxIntersection, ref
//    example_tracer.cpp:18-43) where tn[a] is the exit time the ray had when it last crossed axis a; the exit-time sequence of an
//    axis is monotone (fp32 subtraction and multiplication are monotone), so max_a tn[a] == max(tin, tf[crossed axis]) carried
//    from cell to cell — three registers and the per-axis entry-time selects go away, bit for bit the same value.
//  * the traversal step works on a one-hot crossed-axis mask: the three axes' candidates are formed and the crossed one is kept
//    by selects (an if / else per axis compiled to three divergent 25-instruction paths); the cell coordinates are not carried at
//    all (the exit-plane coordinates pa[] determine them; the lattice is left when pa leaves [0, n/gs]); rare-path state is one
//    flag register.
//  * density lerps: the parity permutation costs one select per lerp instead of two: lerp(p1, p0, t) == p1 + (-t) * (p1 - p0)
//    bit for bit (IEEE negation commutes with rounding), so the difference is formed once and the sign rides on t.
//  * colour: packed nested lerps (fma.rn.f32x2) in slot order with t' = odd ? 1 - t : t (tolerance path), no weights; shaded for
//    zero-weight samples too (a zero weight adds exactly nothing) — no divergent branch, no reconvergence barrier; expf_libm
//    (lrf_device.cuh) is branch-free for the same reason.
// Tie events (edge / corner grazing, exact fp64 compare) re-seed the traversal from scratch like in FastMarch: cold code.
// Measurements and the variants that did not pay: DESIGN.md 4-K1c.
#pragma once
#include "lrf_fast.cuh"

namespace lrf {

#ifndef LRF_LEAN_MINB
#define LRF_LEAN_MINB 3      // resident blocks of 128 threads per SM the register allocation is sized for (168 registers)
#endif
#ifndef LRF_LEAN_NOBRANCH
#define LRF_LEAN_NOBRANCH 1  // 1 = shade the colour of zero-weight samples too instead of branching around it
#endif
#ifndef LRF_LEAN_QPACK
#define LRF_LEAN_QPACK 0     // 1 = the q' lerps packed over z (measured +0.3 %: noise)
#endif
#ifndef LRF_LEAN_PREDADD
#define LRF_LEAN_PREDADD 1   // 1 = node-index updates of the traversal step as predicated adds (12 fewer instructions per sample)
#endif
#ifndef LRF_LEAN_DPACK
#define LRF_LEAN_DPACK 1     // 1 = the x and y density lerps two at a time (add / mul / fma .f32x2: the same IEEE operations)
#endif
#ifndef LRF_LEAN_PREDASM
#define LRF_LEAN_PREDASM 1   // 1 = the pair predicates are formed from the stale mask inside the asm blocks that use them
#endif
#ifndef LRF_LEAN_MERGE
#define LRF_LEAN_MERGE 0     // rare-event regions per sample: 1 = step-length rule and first-cell / tie state share one, 2 = the full window load joins them (spills)
#endif
#ifndef LRF_LEAN_TWOPHASE
#define LRF_LEAN_TWOPHASE 1  // SH chunks requested AFTER the traversal step: 0 = none, 1 = pairs 2, 3 (16 fewer registers in flight across it: no spills,
                             // 2.02 -> 1.975 ms), 2 = pairs 1, 2, 3, 3 = all
#endif

// Predicated gathers of one antipodal slot pair: @p loads come from node A and land in slot J's registers, @!p loads from node B
// and land in slot 7 - J's.  The two addresses are different registers on purpose: with one selected address ptxas merges the
// two predicated loads into one load plus eight predicated moves.
__device__ __forceinline__ void ldg_pair_head(bool p, const float *dA, const float *dB, const float4 *cA, const float4 *cB, float &da, float &db,
                                              float4 &a, float4 &b)
{
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %10, 0;\n\t"
      "@q ld.global.nc.f32 %0, [%11];\n\t@!q ld.global.nc.f32 %1, [%12];\n\t"
      "@q ld.global.nc.v4.f32 {%2, %3, %4, %5}, [%13];\n\t@!q ld.global.nc.v4.f32 {%6, %7, %8, %9}, [%14];\n\t}"
      : "+f"(da), "+f"(db), "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w), "+f"(b.x), "+f"(b.y), "+f"(b.z), "+f"(b.w)
      : "r"((uint32_t)p), "l"(dA), "l"(dB), "l"(cA), "l"(cB));
}
// The same with the predicate formed from the stale mask inside the block (one LOP3 instead of a bool materialised and tested again) and the node
// the SH chunks come from selected on the same predicate.
template <uint32_t BIT>
__device__ __forceinline__ uint32_t ldg_pair_head_m(uint32_t st, uint32_t nA, uint32_t nB, const float *dA, const float *dB, const float4 *cA, const float4 *cB,
                                                    float &da, float &db, float4 &a, float4 &b)
{
  uint32_t nsel;
  asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %11, %12;\n\tsetp.ne.u32 q, t, 0;\n\t"
      "@q ld.global.nc.f32 %0, [%13];\n\t@!q ld.global.nc.f32 %1, [%14];\n\t"
      "@q ld.global.nc.v4.f32 {%2, %3, %4, %5}, [%15];\n\t@!q ld.global.nc.v4.f32 {%6, %7, %8, %9}, [%16];\n\t"
      "selp.u32 %10, %17, %18, q;\n\t}"
      : "+f"(da), "+f"(db), "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w), "+f"(b.x), "+f"(b.y), "+f"(b.z), "+f"(b.w), "=r"(nsel)
      : "r"(st), "n"(BIT), "l"(dA), "l"(dB), "l"(cA), "l"(cB), "r"(nA), "r"(nB));
  return nsel;
}
template <uint32_t BIT> __device__ __forceinline__ void add_sel_m(uint32_t st, float &a, float &b, float x, float y)
{
  asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 q, t, 0;\n\t@q add.rn.f32 %0, %4, %5;\n\t@!q add.rn.f32 %1, %4, %5;\n\t}"
      : "+f"(a), "+f"(b) : "r"(st), "n"(BIT), "f"(x), "f"(y));
}
// @p a = x + y, @!p b = x + y
__device__ __forceinline__ void add_sel(bool p, float &a, float &b, float x, float y)
{
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q add.rn.f32 %0, %3, %4;\n\t@!q add.rn.f32 %1, %3, %4;\n\t}" : "+f"(a), "+f"(b) : "r"((uint32_t)p), "f"(x), "f"(y));
}
// x += v if (m & BIT): one predicate-producing LOP3 and one predicated add (the ?: form compiled to shift / sign-extend / and / add)
template <uint32_t BIT> __device__ __forceinline__ void add_if(uint32_t &x, uint32_t m, uint32_t v)
{
  asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 q, t, 0;\n\t@q add.u32 %0, %0, %3;\n\t}" : "+r"(x) : "r"(m), "n"(BIT), "r"(v));
}
// packed lerp, tolerance path: a + t * (b - a) as two fma.rn.f32x2
__device__ __forceinline__ u64 lerp2(u64 a, u64 b, u64 t, u64 neg1) { return fma2(t, fma2(a, neg1, b), a); }

// exact comparison of the three exit-plane times (the reference's own fp64 quotients); returns the crossing mask.  Cold, out of line.
static __device__ __noinline__ int lean_exact_cross(float p0, float p1, float p2, float o0, float o1, float o2, float d0, float d1, float d2)
{
  const double t0 = ((double)p0 - (double)o0) / (double)d0, t1 = ((double)p1 - (double)o1) / (double)d1, t2 = ((double)p2 - (double)o2) / (double)d2;
  const double em = fmin(t0, fmin(t1, t2));
  return (t0 == em ? 1 : 0) | (t1 == em ? 2 : 0) | (t2 == em ? 4 : 0);
}

// GATE = false compiles the step-length rule of ref :145-146 out (experiment only, see launch_render_image: it is NOT safe behind tie events)
template <bool TAPS, bool GATE = true> struct LeanMarch
{
  static constexpr uint32_t kFull = 1u, kResync = 2u, kLeaving = 4u, kSeeded = 8u, kHit = 16u, kQueue = 0xff00u;

  const SceneDev &S;
  const uint64_t *etab;
  float o_[3], d_[3], idir_[3];
  float tf[3], pa[3];            // exit time / exit-plane coordinate of the current cell, per axis
  float tin;                     // == max_a tn[a] of the current cell (see the header)
  uint32_t nx[2], ny[2], nz[2];  // node index contribution held by parity p of each axis: x, y*gs, z*gs^2
  uint32_t odd, negmask;         // parity bits of the cell coordinates; axes the ray travels down
  uint32_t stale = 0;            // bit J (J = 0..3): slot J is the stale member of pair (J, 7 - J), else slot 7 - J (or none)
  uint32_t cold = 0;             // rare-path flags, one register tested on the hot path: kFull | kResync | kLeaving | kSeeded, queue length << 8
  float hs[3];                   // per axis: offset from the exit plane back to the cell's lower plane (1/gs or 0)
  uint32_t dn2[3];               // per axis: signed node-index step of two nodes
  float dn[8];
  float4 cw[8];                  // {R, G, B, sh0} as loaded
  float qr[8];                   // sum_{j >= 1} sh_j Y_j(d): the part of q = sh . Y(d) beyond sh0 * Y0 (kept apart so that nothing touches cw before the shading)
  float Y0;
  u64 Y12, Y34, Y56, Y78;
  float thr = 1.0f, tPrev = -1.0f;
  float col[3] = {0.0f, 0.0f, 0.0f};
  uint32_t n_shaded = 0;
  __device__ __forceinline__ bool any_hit() const { return (cold & kHit) != 0; }   // set with the first cell (on a dense lattice the first event always holds at least one cell)

  __device__ LeanMarch(const SceneDev &s, const uint64_t *t) : S(s), etab(t) {}
  __device__ __forceinline__ uint32_t slot_node(int s) const { return nx[s & 1] + ny[(s >> 1) & 1] + nz[s >> 2]; }
  __device__ __forceinline__ int sgn(int a) const { return (negmask >> a) & 1u ? -1 : 1; }
  // cell coordinate along a, from the exit plane (cold paths only)
  __device__ __forceinline__ int cell_of(int a) const { return (int)(pa[a] * S.gs_f) - ((negmask >> a) & 1u ? 0 : 1); }

  __device__ __forceinline__ float contract(const float4 &b, const float4 &e) const
  {
    u64 q2 = mul2(pk2(b.x, b.y), Y12);
    q2 = fma2(pk2(b.z, b.w), Y34, q2);
    q2 = fma2(pk2(e.x, e.y), Y56, q2);
    q2 = fma2(pk2(e.z, e.w), Y78, q2);
    float ql, qh;
    upk2(q2, ql, qh);
    return ql + qh;
  }
  // cold: every piece of per-cell state from scratch
  __device__ __forceinline__ void set_cell(int x, int y, int z)
  {
    const int v[3] = {x, y, z};
    uint32_t nd[3][2];
    float tn[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const bool pos = !((negmask >> a) & 1u);
      const int ex = v[a] + (pos ? 1 : 0), en = v[a] + (pos ? 0 : 1);
      pa[a] = (float)ex * S.inv_gs_f;
      tf[a] = idir_[a] * (pa[a] - o_[a]);
      tn[a] = idir_[a] * ((float)en * S.inv_gs_f - o_[a]);
      nd[a][0] = (uint32_t)(v[a] + (v[a] & 1));
      nd[a][1] = (uint32_t)(v[a] + 1 - (v[a] & 1));
    }
    tin = fmaxf(fmaxf(tn[0], tn[1]), tn[2]);
    const uint32_t gs = S.gs, gs2 = (uint32_t)S.gs * S.gs;
    nx[0] = nd[0][0]; nx[1] = nd[0][1];
    ny[0] = nd[1][0] * gs; ny[1] = nd[1][1] * gs;
    nz[0] = nd[2][0] * gs2; nz[1] = nd[2][1] * gs2;
    odd = (uint32_t)(x & 1) | ((uint32_t)(y & 1) << 1) | ((uint32_t)(z & 1) << 2);
    cold |= kFull;
  }
  // cold: blocking load of all 8 window slots
  __device__ __forceinline__ void load_all()
  {
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const uint32_t node = slot_node(s);
      const float4 *pc = S.col + (size_t)node * kColF4;
      const float4 a = __ldg(pc), b = __ldg(pc + 1), e = __ldg(pc + 2);
      dn[s] = __ldg(S.dens + node);
      cw[s] = a;
      qr[s] = contract(b, e);
    }
    cold &= ~kFull;
    stale = 0;
  }

  struct Geom
  {
    float delta, lf[3];
    uint32_t odd;
  };
  struct Pending
  {
    uint32_t stale;
    float4 b[4], e[4];
    Geom g;
  };
  template <int J> __device__ __forceinline__ const float4 *issue_pair(uint32_t st)
  {
    const uint32_t nA = slot_node(J), nB = slot_node(7 - J);
#if LRF_LEAN_PREDASM
    const uint32_t ns = ldg_pair_head_m<(1u << J)>(st, nA, nB, S.dens + nA, S.dens + nB, S.col + (size_t)nA * kColF4, S.col + (size_t)nB * kColF4, dn[J], dn[7 - J],
                                                   cw[J], cw[7 - J]);
    return S.col + (size_t)ns * kColF4;
#else
    const bool p = (st & (1u << J)) != 0;
    ldg_pair_head(p, S.dens + nA, S.dens + nB, S.col + (size_t)nA * kColF4, S.col + (size_t)nB * kColF4, dn[J], dn[7 - J], cw[J], cw[7 - J]);
    return S.col + (size_t)(p ? nA : nB) * kColF4;             // the SH chunks come from ONE address: an unsplit load costs fewer L1 wavefronts
#endif
  }
  template <int J> __device__ __forceinline__ void commit_pair(uint32_t st, const float4 &b, const float4 &e)
  {
    u64 q2 = mul2(pk2(b.x, b.y), Y12);
    q2 = fma2(pk2(b.z, b.w), Y34, q2);
    q2 = fma2(pk2(e.x, e.y), Y56, q2);
    q2 = fma2(pk2(e.z, e.w), Y78, q2);
    float ql, qh;
    upk2(q2, ql, qh);
#if LRF_LEAN_PREDASM
    add_sel_m<(1u << J)>(st, qr[J], qr[7 - J], ql, qh);
#else
    add_sel((st & (1u << J)) != 0, qr[J], qr[7 - J], ql, qh);
#endif
  }
  // geometry of the sample in the current cell, ref RayGridIntersection example_tracer.cpp:134-153 (bit-exact)
  template <bool G> __device__ __forceinline__ void geometry(Geom &g)
  {
    const float tmax = fminf(fminf(tf[0], tf[1]), tf[2]);      // all finite: == the std::min chain of :18-43
    if (tPrev < 0.0f) tPrev = tin;                             // :139
    const float t = (tin + tmax) * 0.5f;                       // :142
    float delta = t - tPrev;                                   // :144
#if !LRF_LEAN_MERGE
    if (G && delta > S.step_gate_f)                            // :145-146
      delta = (float)((double)delta - (double)((int)((double)delta / S.step) - 1) * S.step);
#endif
    tPrev = t;                                                 // :190
    g.delta = delta;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float bmin = pa[a] - hs[a];                        // == (float)cell / (float)gs, exact
      const float p = o_[a] + t * d_[a];                       // :148
      g.lf[a] = (p - bmin) * S.gs_f;                           // :150
    }
    g.odd = odd;
  }
  // one sample from the window; false = the reference's throughput break fired (:131)
  __device__ __forceinline__ bool shade(const Geom &g)
  {
    const bool od0 = (g.odd & 1u) != 0, od1 = (g.odd & 2u) != 0, od2 = (g.odd & 4u) != 0;
    // ---- density: the reference's nested lerps (:155-167 via lerpCell :72-80), bit-exact ----
    const float s0 = od0 ? -g.lf[0] : g.lf[0], s1 = od1 ? -g.lf[1] : g.lf[1], s2 = od2 ? -g.lf[2] : g.lf[2];
    float B[2];
#if LRF_LEAN_DPACK
    {   // the same IEEE operations, the subtraction and the addition two at a time: p1 - p0 as fma(p0, -1, p1) and base + m as fma(m, 1, base)
        // (exact products).  The multiplication stays scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 — and even mul.rn.f32x2 +
        // fma.rn.f32x2(m, 1, c) — into one FFMA2, which rounds once where the reference rounds twice.
      const u64 N1 = pk2(-1.0f, -1.0f), P1 = pk2(1.0f, 1.0f);
      auto lerp_pair = [&](u64 p0, u64 p1, float sg, bool od) {
        float dl, dh;
        upk2(fma2(p0, N1, p1), dl, dh);
        return fma2(pk2(sg * dl, sg * dh), P1, od ? p1 : p0);
      };
      const u64 Aa = lerp_pair(pk2(dn[0], dn[4]), pk2(dn[1], dn[5]), s0, od0);   // {A[0], A[2]}
      const u64 Ab = lerp_pair(pk2(dn[2], dn[6]), pk2(dn[3], dn[7]), s0, od0);   // {A[1], A[3]}
      upk2(lerp_pair(Aa, Ab, s1, od1), B[0], B[1]);
    }
#else
    float A[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float p0 = dn[2 * j], p1 = dn[2 * j + 1];
      A[j] = (od0 ? p1 : p0) + s0 * (p1 - p0);
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const float p0 = A[2 * j], p1 = A[2 * j + 1];
      B[j] = (od1 ? p1 : p0) + s1 * (p1 - p0);
    }
#endif
    float sigma = (od2 ? B[1] : B[0]) + s2 * (B[1] - B[0]);
    if (sigma < 0.0f) sigma = 0.0f;                            // :170-171
    const float tr = expf_libm(-sigma * g.delta, etab);        // :175 (branch-free, lrf_device.cuh)
    const float wgt = thr * (1.0f - tr);                       // :187
#if LRF_LEAN_NOBRANCH
    {   // colour always (a zero weight adds exactly nothing): no divergent branch, no reconvergence barrier
#else
    if (wgt != 0.0f) {
#endif
      // ---- colour: 1e-4-tolerance path, packed nested lerps in slot order ----
      const float tx = od0 ? 1.0f - g.lf[0] : g.lf[0], ty = od1 ? 1.0f - g.lf[1] : g.lf[1], tz = od2 ? 1.0f - g.lf[2] : g.lf[2];
      const u64 TX = pk2(tx, tx), TY = pk2(ty, ty), TZ = pk2(tz, tz), N1 = pk2(-1.0f, -1.0f);
      u64 rg[4], bs[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        rg[j] = lerp2(pk2(cw[2 * j].x, cw[2 * j].y), pk2(cw[2 * j + 1].x, cw[2 * j + 1].y), TX, N1);
        bs[j] = lerp2(pk2(cw[2 * j].z, cw[2 * j].w), pk2(cw[2 * j + 1].z, cw[2 * j + 1].w), TX, N1);
      }
      const u64 rg0 = lerp2(rg[0], rg[1], TY, N1), rg1 = lerp2(rg[2], rg[3], TY, N1);
      const u64 bs0 = lerp2(bs[0], bs[1], TY, N1), bs1 = lerp2(bs[2], bs[3], TY, N1);
#if LRF_LEAN_QPACK
      // q' lerps packed over z: (slot s, slot s + 4) pairs go through the x and y lerps together, the z lerp joins the halves
      const u64 qx0 = lerp2(pk2(qr[0], qr[4]), pk2(qr[1], qr[5]), TX, N1), qx1 = lerp2(pk2(qr[2], qr[6]), pk2(qr[3], qr[7]), TX, N1);
      float qb0, qb1;
      upk2(lerp2(qx0, qx1, TY, N1), qb0, qb1);
#else
      float qa[4];
#pragma unroll
      for (int j = 0; j < 4; j++) qa[j] = fmaf(tx, qr[2 * j + 1] - qr[2 * j], qr[2 * j]);
      const float qb0 = fmaf(ty, qa[1] - qa[0], qa[0]), qb1 = fmaf(ty, qa[3] - qa[2], qa[2]);
#endif
      float R, G, Bc, S0;
      upk2(lerp2(rg0, rg1, TZ, N1), R, G);
      upk2(lerp2(bs0, bs1, TZ, N1), Bc, S0);
      const float Q = fmaf(S0, Y0, fmaf(tz, qb1 - qb0, qb0));
      const float s = __saturatef(Q);                          // :181
      col[0] = fmaf(wgt, __saturatef(R * s), col[0]);          // :182-187
      col[1] = fmaf(wgt, __saturatef(G * s), col[1]);
      col[2] = fmaf(wgt, __saturatef(Bc * s), col[2]);
    }
    thr *= tr;                                                 // :189
    if (TAPS) n_shaded++;
    return !(thr < 0.1f);                                      // :131
  }
  __device__ void run(const float *of, const float *df)
  {
    negmask = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      o_[a] = of[a];
      d_[a] = df[a];
      idir_[a] = 1.0f / df[a];                                 // ref example_tracer.cpp:20-22
      negmask |= df[a] > 0.0f ? 0u : (1u << a);
      hs[a] = df[a] > 0.0f ? S.inv_gs_f : 0.0f;
    }
    {
      const uint32_t gs = (uint32_t)S.gs;
      dn2[0] = (negmask & 1u) ? 0u - 2u : 2u;
      dn2[1] = (negmask & 2u) ? 0u - 2u * gs : 2u * gs;
      dn2[2] = (negmask & 4u) ? 0u - 2u * gs * gs : 2u * gs * gs;
    }
    {
      float Y[9];
      sh_eval_2(df, Y);
      Y0 = Y[0];
      Y12 = pk2(Y[1], Y[2]); Y34 = pk2(Y[3], Y[4]); Y56 = pk2(Y[5], Y[6]); Y78 = pk2(Y[7], Y[8]);
    }
    const int n = S.n;
    // ---- entry: exact fp64, once per ray (the construction of FastMarch::run / march) ----
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
        t_enter = fmax(t_enter, sgn(a) > 0 ? T0 : Tn);
        t_exit = fmin(t_exit, sgn(a) > 0 ? Tn : T0);
      }
      const double t_start = fmax(t_enter, 0.0);
      if (!(t_exit >= t_start)) return;                        // ref octree.h:483 for the whole lattice
      int lo[3], two_mask = 0;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const int s = sgn(a);
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
    const int pos_mask = (int)(~negmask & 7u);
    const float pmax = (float)n * S.inv_gs_f;                  // exact
    cold = kResync | (leaving_after ? kLeaving : 0u) | ((uint32_t)nq << 8);

    Pending P;
    // cold: first cell, tie events.  0 = go on to the face selection, 1 = the next cell is set, 2 = the ray is done
    auto cold_part = [&]() -> int {
      int x, y, z;
      const bool from_queue = (cold & kQueue) != 0;
      bool reseed = true;
      if (from_queue) {
        cold -= 0x100u;
        unpack_cell(queue[cold >> 8], x, y, z);
      } else {
        if (cold & kLeaving) return 2;
        x = cur[0]; y = cur[1]; z = cur[2];
        reseed = !(cold & kSeeded) || cell_of(0) != x || cell_of(1) != y || cell_of(2) != z;
        cold &= ~kResync;
      }
      if (reseed) set_cell(x, y, z);
      cold |= kSeeded;
      if (from_queue) {
        cold |= kHit;
        return 1;
      }
      return 0;
    };
    // moves the traversal to the next cell to be shaded; false = the ray is done
    auto next_cell = [&]() -> bool {
#if LRF_LEAN_MERGE
      // ONE rare region per sample in front of the face selection: the full window load a re-seed asked for (the stale-pair requests issued
      // just before it reload slots from their own nodes: harmless), the step-length rule (ref :145-146), and the first-cell / tie-event state
      if ((cold & ((LRF_LEAN_MERGE > 1 ? kFull : 0u) | kQueue | kResync)) | (uint32_t)(GATE && P.g.delta > S.step_gate_f)) {
        if (LRF_LEAN_MERGE > 1 && (cold & kFull)) load_all();
        if (GATE && P.g.delta > S.step_gate_f)
          P.g.delta = (float)((double)P.g.delta - (double)((int)((double)P.g.delta / S.step) - 1) * S.step);
        if (cold & (kQueue | kResync)) {
          const int r = cold_part();
          if (r) return r == 1;
        }
      }
#endif
      for (;;) {
#if !LRF_LEAN_MERGE
        if (cold & (kQueue | kResync)) {
          const int r = cold_part();
          if (r) return r == 1;
        }
#endif
        // ---- which face does the ray leave through?  fp32 filter, exact fp64 when close (lrf_fast.cuh) ----
        const float m = fminf(fminf(tf[0], tf[1]), tf[2]);
        const float gate = fmaf(fabsf(m), 0x1p-20f, 0x1p-100f);
        const bool c0 = tf[0] - m <= gate, c1 = tf[1] - m <= gate, c2 = tf[2] - m <= gate;
        float tcross = m;
        uint32_t cm = c0 ? 1u : (c1 ? 2u : 4u);                // crossed axis, one-hot
        if ((c0 && (c1 || c2)) || (c1 && c2)) {                // cold (about 1 step in 1000)
          const int cross = lean_exact_cross(pa[0], pa[1], pa[2], o_[0], o_[1], o_[2], d_[0], d_[1], d_[2]);
          if (cross & (cross - 1)) {
            // true tie: the ray passes through a lattice edge / corner — 3 or 7 boxes at the same distance
            int lo[3];
            cold &= ~kLeaving;
#pragma unroll
            for (int a = 0; a < 3; a++) {
              const int c = cell_of(a);
              lo[a] = c;
              cur[a] = c;
              if (cross & (1 << a)) {
                const int nw = c + sgn(a);
                lo[a] = min(c, nw);
                cur[a] = nw;
                if ((unsigned)nw >= (unsigned)n) cold |= kLeaving;
              }
            }
            uint32_t cells[8];
            const int mm = build_event_cells(S, lo[0], lo[1], lo[2], cross, cross, pos_mask, cells);
            for (int j = 0; j < mm; j++) queue[j] = cells[mm - 1 - j];
            cold = (cold & ~kQueue) | ((uint32_t)mm << 8) | kResync;
#if LRF_LEAN_MERGE
            {
              const int r = cold_part();
              if (r) return r == 1;
            }
#endif
            continue;
          }
          cm = (uint32_t)cross;
          tcross = (cross & 1) ? tf[0] : ((cross & 2) ? tf[1] : tf[2]);
        }
        // ---- one face crossing, branch-free: the three axes' candidates are formed and the crossed one is kept ----
        tin = fmaxf(tin, tcross);
        const uint32_t pl = odd ^ negmask;                     // parity of the node plane that leaves the window, per axis
        odd ^= cm;
        const uint32_t up0 = cm & ~pl, up1 = cm & pl;          // axis whose parity-0 / parity-1 node plane is replaced
        const bool x0 = (cm & 1u) != 0, x1 = (cm & 2u) != 0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const bool on = (cm & (1u << a)) != 0;
          // signed cell pitch copysign(1/gs, d): one LOP3 on the spot (as a per-ray register it was spilled and re-read every iteration)
          const float sda = __uint_as_float(__float_as_uint(S.inv_gs_f) | (__float_as_uint(d_[a]) & 0x80000000u));
          const float pan = pa[a] + sda;                       // exact: multiples of 2^-k
          const float tfn = idir_[a] * (pan - o_[a]);
          pa[a] = on ? pan : pa[a];
          tf[a] = on ? tfn : tf[a];
        }
        const float pnew = x0 ? pa[0] : (x1 ? pa[1] : pa[2]);
#if LRF_LEAN_PREDADD
        add_if<1>(nx[0], up0, dn2[0]); add_if<1>(nx[1], up1, dn2[0]);
        add_if<2>(ny[0], up0, dn2[1]); add_if<2>(ny[1], up1, dn2[1]);
        add_if<4>(nz[0], up0, dn2[2]); add_if<4>(nz[1], up1, dn2[2]);
#else
        nx[0] += (up0 & 1u) ? dn2[0] : 0u; nx[1] += (up1 & 1u) ? dn2[0] : 0u;
        ny[0] += (up0 & 2u) ? dn2[1] : 0u; ny[1] += (up1 & 2u) ? dn2[1] : 0u;
        nz[0] += (up0 & 4u) ? dn2[2] : 0u; nz[1] += (up1 & 4u) ? dn2[2] : 0u;
#endif
        // stale low members: the slots J < 4 whose bit along the crossed axis equals the leaving parity (x: 5 / a, y: 3 / c, z: f / 0)
        stale = ((0xf35u >> ((cm >> 1) * 4u)) & 0xfu) ^ (up1 ? 0xfu : 0u);
        if (pnew > pmax || pnew < 0.0f) return false;          // left the lattice
        return true;
      }
    };

    P.g.delta = 0.0f;
    if (!next_cell()) return;
    for (;;) {
      // One sample: (1) request the stale half of the window for the current cell, (2) the cell's geometry and the whole traversal
      // step to the next cell while the loads fly, (3) contract, shade, composite.
#if LRF_LEAN_MERGE < 2
      if (cold & kFull) load_all();
#endif
      P.stale = stale;
      const float4 *c0 = issue_pair<0>(P.stale), *c1 = issue_pair<1>(P.stale), *c2 = issue_pair<2>(P.stale), *c3 = issue_pair<3>(P.stale);
      // the SH chunks of the first (4 - LRF_LEAN_TWOPHASE * 2 ...) pairs fly across the traversal step, the rest are requested behind it
      if (LRF_LEAN_TWOPHASE < 3) { P.b[0] = __ldg(c0 + 1); P.e[0] = __ldg(c0 + 2); }
      if (LRF_LEAN_TWOPHASE < 2) { P.b[1] = __ldg(c1 + 1); P.e[1] = __ldg(c1 + 2); }
      if (LRF_LEAN_TWOPHASE < 1) {
        P.b[2] = __ldg(c2 + 1); P.e[2] = __ldg(c2 + 2);
        P.b[3] = __ldg(c3 + 1); P.e[3] = __ldg(c3 + 2);
      }
      geometry<GATE>(P.g);
      const bool more = next_cell();
      if (LRF_LEAN_TWOPHASE >= 3) { P.b[0] = __ldg(c0 + 1); P.e[0] = __ldg(c0 + 2); }
      if (LRF_LEAN_TWOPHASE >= 2) { P.b[1] = __ldg(c1 + 1); P.e[1] = __ldg(c1 + 2); }
      if (LRF_LEAN_TWOPHASE >= 1) {
        P.b[2] = __ldg(c2 + 1); P.e[2] = __ldg(c2 + 2);
        P.b[3] = __ldg(c3 + 1); P.e[3] = __ldg(c3 + 2);
      }
      commit_pair<0>(P.stale, P.b[0], P.e[0]); commit_pair<1>(P.stale, P.b[1], P.e[1]);
      commit_pair<2>(P.stale, P.b[2], P.e[2]); commit_pair<3>(P.stale, P.b[3], P.e[3]);
      if (!shade(P.g)) return;
      if (!more) return;
    }
  }
};

}  // namespace lrf
