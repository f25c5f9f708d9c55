// lrf_wray.cuh — warp-per-ray, sample-parallel backward for packed ray batches (BASELINE config 3: 4096 random rays
// per optimisation step).
//
// With a few thousand incoherent rays there are fewer rays than the GPU has warp slots (148 SMs x 64), and a
// thread-per-ray kernel leaves each ray on one lane with ~500 dependent steps of gathers: the first version ran the
// 4096-ray batch in 1.8 ms at 6 % warp occupancy and 10 % issue utilisation.  Here one WARP takes one ray:
//   * the exact traversal (same fp32-filtered DDA as lrf_fast.cuh) runs in lock-step on all lanes — it is the only
//     serial part — and lane k latches the k-th box of each group of 32 hits;
//   * every lane then shades ITS sample (8-corner gathers, trilinear density, libm-exact expf, SH colour) in parallel;
//   * the throughput product — which must be the reference's left-to-right fp32 product, because the break at
//     `throughput < 0.1` (ref example_tracer.cpp:131) decides how many samples exist — is formed by a 32-step
//     shuffle loop, a few instructions per sample;
//   * the adjoint of sample k needs only prefix/suffix sums of the composited colour (closed form of the reverse sweep,
//     see grad_kernel), so the backward is sample-parallel too: a warp scan, then each lane scatters its 8 x 13 values
//     with red.global.add.{f32,v4.f32}.
// Pass A (forward) walks the ray and leaves its hit list — 12 bytes per hit: packed cell + the two fp32 slab times — in a
// per-ray scratch slice; pass B (adjoint + scatter) reads the list back instead of walking again (the lock-step traversal is
// ~3/4 of the kernel's instructions).  A ray whose list outgrows the slice simply walks twice.
// ncu (profiles/r1n_wray_ncu_summary.txt): 27 K warp instructions per ray, issue slots 26 % busy, 51 % of the stall samples on
// the load scoreboard, 10 warps per SM — the gathers' latency is the bound.  Tried without gain: 128 / 96 registers per thread
// (16 / 20 warps per SM: -2.5 % / +16 %) and an L2 prefetch of the next group's nodes from the recorded list (+2 %).
#pragma once
#include "lrf_fast.cuh"

#ifndef LRF_WRAY_PARALLEL
#define LRF_WRAY_PARALLEL 1   // dense sets: hits of a group computed by all lanes at once (gather_parallel); 0 = always the lock-step walk
#endif
#ifndef LRF_WRAY_PREFETCH
#define LRF_WRAY_PREFETCH 0   // colour lines of a group's corners requested while its density is being shaded
#endif
#ifndef LRF_WRAY_CHECK
#define LRF_WRAY_CHECK 0      // debug build: run both and print mismatches
#endif

namespace lrf {

// Exact hit enumeration without any node window: the traversal half of FastMarch.
template <bool SPARSE> struct RayWalker
{
  const SceneDev &S;
  float o[3], d[3], idir[3];
  int cell[3], sgn[3];
  float pa[3], tn[3], tf[3];
  uint32_t queue[8];
  int nq = 0, cur[3], pos_mask = 0;
  bool leaving_after = false, resync = true;

  __device__ RayWalker(const SceneDev &s) : S(s) {}
  // the mutable part of the walker (debug builds compare the lane-parallel enumeration with the lock-step walk)
  struct State
  {
    int cell[3], cur[3], nq;
    float pa[3], tn[3], tf[3];
    uint32_t queue[8];
    bool leaving_after, resync;
  };
  __device__ __forceinline__ State save() const
  {
    State t;
    for (int a = 0; a < 3; a++) { t.cell[a] = cell[a]; t.cur[a] = cur[a]; t.pa[a] = pa[a]; t.tn[a] = tn[a]; t.tf[a] = tf[a]; }
    for (int i = 0; i < 8; i++) t.queue[i] = queue[i];
    t.nq = nq; t.leaving_after = leaving_after; t.resync = resync;
    return t;
  }
  __device__ __forceinline__ void restore(const State &t)
  {
    for (int a = 0; a < 3; a++) { cell[a] = t.cell[a]; cur[a] = t.cur[a]; pa[a] = t.pa[a]; tn[a] = t.tn[a]; tf[a] = t.tf[a]; }
    for (int i = 0; i < 8; i++) queue[i] = t.queue[i];
    nq = t.nq; leaving_after = t.leaving_after; resync = t.resync;
  }

  __device__ __forceinline__ void set_cell(int x, int y, int z)
  {
    const int v[3] = {x, y, z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      cell[a] = v[a];
      const int ex = v[a] + (sgn[a] > 0 ? 1 : 0), en = v[a] + (sgn[a] > 0 ? 0 : 1);
      pa[a] = (float)ex * S.inv_gs_f;
      tf[a] = idir[a] * (pa[a] - o[a]);
      tn[a] = idir[a] * ((float)en * S.inv_gs_f - o[a]);
    }
  }
  __device__ __forceinline__ int advance(int ax)
  {
    const bool mm[3] = {ax == 0, ax == 1, ax == 2};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float pan = pa[a] + (sgn[a] > 0 ? S.inv_gs_f : -S.inv_gs_f);
      const float tfn = idir[a] * (pan - o[a]);
      tn[a] = mm[a] ? tf[a] : tn[a];
      tf[a] = mm[a] ? tfn : tf[a];
      pa[a] = mm[a] ? pan : pa[a];
      cell[a] = mm[a] ? cell[a] + sgn[a] : cell[a];
    }
    return mm[0] ? cell[0] : (mm[1] ? cell[1] : cell[2]);
  }
  __device__ __forceinline__ int decide_fast() const
  {
    const float m = fminf(fminf(tf[0], tf[1]), tf[2]);
    const float gate = fmaf(fabsf(m), 0x1p-20f, 0x1p-100f);
    const bool c0 = tf[0] - m <= gate, c1 = tf[1] - m <= gate, c2 = tf[2] - m <= gate;
    const int ax = c0 ? 0 : (c1 ? 1 : 2);
    return ((c0 && c1) || (c0 && c2) || (c1 && c2)) ? -1 : ax;
  }
  __device__ __forceinline__ int exact_cross() const
  {
    double te[3];
#pragma unroll
    for (int a = 0; a < 3; a++) te[a] = ((double)pa[a] - (double)o[a]) / (double)d[a];
    const double em = fmin(te[0], fmin(te[1], te[2]));
    return (te[0] == em ? 1 : 0) | (te[1] == em ? 2 : 0) | (te[2] == em ? 4 : 0);
  }

  // entry into the lattice (exact fp64, same construction as the generic march); false = the ray misses it
  __device__ __forceinline__ bool start(const float *of, const float *df)
  {
    const int n = S.n;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      o[a] = of[a];
      d[a] = df[a];
      idir[a] = 1.0f / df[a];
      sgn[a] = df[a] > 0.0f ? 1 : -1;
    }
    nq = 0;
    leaving_after = false;
    resync = true;
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
    if (!(t_exit >= t_start)) return false;
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
      for (int j = 0; j < m; j++) queue[j] = cells[m - 1 - j];
      nq = m;
    }
    pos_mask = (sgn[0] > 0 ? 1 : 0) | (sgn[1] > 0 ? 2 : 0) | (sgn[2] > 0 ? 4 : 0);
    cell[0] = cell[1] = cell[2] = -8;
    return true;
  }

  // moves to the next hit; false = no more hits
  __device__ __forceinline__ bool next()
  {
    const int n = S.n;
    for (;;) {
      if (nq > 0) {
        int x, y, z;
        unpack_cell(queue[--nq], x, y, z);
        set_cell(x, y, z);
        return true;
      }
      if (resync) {
        if (leaving_after) return false;
        if (cell[0] != cur[0] || cell[1] != cur[1] || cell[2] != cur[2]) set_cell(cur[0], cur[1], cur[2]);
        resync = false;
      }
      int ax = decide_fast();
      if (ax < 0) {
        const int cross = exact_cross();
        if (cross & (cross - 1)) {
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
          continue;
        }
        ax = (cross & 1) ? 0 : ((cross & 2) ? 1 : 2);
      }
      const int moved = advance(ax);
      if ((unsigned)moved >= (unsigned)n) return false;
      if (SPARSE && !cell_present(S, cell[0], cell[1], cell[2])) continue;
      return true;
    }
  }
};

__device__ __forceinline__ void red_add_f32(float *p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add_v4(float4 *p, float a, float b, float c, float d)
{
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += y;
  }
  return v;
}

// hit list entry left by pass A for pass B
struct HitRec
{
  uint32_t cell;      // pack_cell
  float tmin, tmax;
};

// one lane's sample of a 32-hit group
struct LaneSample
{
  int x, y, z;
  float tmin, tmax;
  // derived
  float delta, lf[3], w[8];
  uint32_t base;
  float dens_raw, sigma, tr, T;   // T = throughput BEFORE this sample
  float rgb[3], raw, s, c[3];     // interpolated RGB, sh . Y, clamped s, clamped colour
};

template <bool SPARSE, bool GREY> struct WarpRay
{
  const SceneDev &S;
  const uint64_t *etab;
  const int lane;
  RayWalker<SPARSE> W;
  float Y[9];
  bool more = false;     // the walker currently stands on an unconsumed hit

  uint32_t *slot = nullptr;   // this warp's 36 words of shared memory for gather_parallel (nullptr: always walk serially)

  __device__ WarpRay(const SceneDev &s, const uint64_t *t, uint32_t *sl = nullptr) : S(s), etab(t), lane(threadIdx.x & 31), W(s), slot(sl) {}

  __device__ __forceinline__ bool begin(const float *o, const float *d)
  {
    if (!W.start(o, d)) return false;
    more = W.next();
    return more;
  }
  // Dense sets: the next `want` hits by all lanes AT ONCE instead of 32 lock-step traversal steps (which were ~3/4 of this kernel's
  // instructions).  On a dense lattice the hit sequence is the merge of three monotone sequences: the j-th crossing of axis a happens
  // at  t_a(j) = idir_a * (plane_a(j) / gs - o_a)  — the very fp32 value the serial walk carries as tf[a] — and the r-th crossing
  // overall leads to hit r.  Lane l takes crossing j = l + 1 of every axis, COUNTS the crossings of the other two axes that come
  // before it (an estimate from the pitch, corrected against the exact fp32 times), and so knows its rank r = j + count_b + count_c
  // and the cell it leads to: c0 + sgn * (n_x, n_y, n_z).  It leaves the triple in slot[r]; lane r picks it up and forms the hit's
  // slab times from t_a(n_a) / t_a(n_a + 1).  Whenever two crossing times of the first 33 are within the rounding bound of each
  // other (the serial walk would consult the exact fp64 quotients, and a true tie expands into 3 / 7 boxes) the group is walked
  // serially instead: returns -1 with nothing changed.  Otherwise the fp32 order IS the exact order (lrf_fast.cuh) and the result is
  // the serial walk's, bit for bit.
  __device__ __forceinline__ int gather_parallel(LaneSample &L, int want)
  {
    if (SPARSE || slot == nullptr || !more || W.nq != 0 || W.leaving_after) return -1;
    if (W.resync && (W.cell[0] != W.cur[0] || W.cell[1] != W.cur[1] || W.cell[2] != W.cur[2])) return -1;
    const int n = S.n;
    const int c0[3] = {W.cell[0], W.cell[1], W.cell[2]};
    const int E[3] = {c0[0] + (W.sgn[0] > 0 ? 1 : 0), c0[1] + (W.sgn[1] > 0 ? 1 : 0), c0[2] + (W.sgn[2] > 0 ? 1 : 0)};
    auto tcross = [&](int a, int j) { return W.idir[a] * ((float)(E[a] + (j - 1) * W.sgn[a]) * S.inv_gs_f - W.o[a]); };   // exact plane coordinate: gs = 2^k
    const float INF = __int_as_float(0x7f800000);
    __syncwarp();                                              // the previous group's slots have been read
    const int j = lane + 1;
    bool amb = false;
    uint32_t my_exit = 99u;
    if (lane == 0) slot[0] = 0u;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float T = tcross(a, j);
      const float gate = fmaf(fabsf(T), 0x1p-19f, 0x1p-90f);
      int cnt[3];
      bool amb_a = false;
      cnt[a] = j;
#pragma unroll
      for (int q = 1; q <= 2; q++) {
        const int b = (a + q) % 3;
        // crossings of axis b strictly before T: t_b(k) < T <= t_b(k + 1)
        // (the estimate from the pitch is off by at most one; one correction step, then the bracket is CHECKED — a miss walks serially)
        int k = 1 + (int)floorf((T - W.tf[b]) * (fabsf(W.d[b]) * S.gs_f));
        k = max(0, min(k, 40));
        float lo = k >= 1 ? tcross(b, k) : -INF, hi = tcross(b, k + 1);
        if (!(lo < T)) { k--; hi = lo; lo = k >= 1 ? tcross(b, k) : -INF; }
        else if (hi < T) { k++; lo = hi; hi = tcross(b, k + 1); }
        if (k < 0 || !(lo < T) || hi < T || T - lo <= gate || hi - T <= gate) amb_a = true;
        cnt[b] = max(k, 0);
      }
      const int rank = cnt[0] + cnt[1] + cnt[2];
      if (rank <= 32) {
        const bool exits = (unsigned)(c0[a] + j * W.sgn[a]) >= (unsigned)n;
        slot[rank] = (uint32_t)cnt[0] | ((uint32_t)cnt[1] << 8) | ((uint32_t)cnt[2] << 16);
        if (exits) my_exit = min(my_exit, (uint32_t)rank);
        amb = amb || amb_a;
      }
    }
    if (__any_sync(0xffffffffu, amb)) return -1;
    const int r_exit = (int)__reduce_min_sync(0xffffffffu, my_exit);   // the crossing of this rank leaves the lattice: hits are ranks 0 .. r_exit - 1
    const int m = min(want, r_exit);
    __syncwarp();
    auto times = [&](uint32_t packed, int *nn, float *tn_, float *tf_) {
      nn[0] = (int)(packed & 255u); nn[1] = (int)((packed >> 8) & 255u); nn[2] = (int)((packed >> 16) & 255u);
#pragma unroll
      for (int a = 0; a < 3; a++) {
        tn_[a] = nn[a] >= 1 ? tcross(a, nn[a]) : W.tn[a];
        tf_[a] = tcross(a, nn[a] + 1);
      }
    };
    if (lane < m) {
      int nn[3];
      float tn_[3], tf_[3];
      times(slot[lane], nn, tn_, tf_);
      L.x = c0[0] + nn[0] * W.sgn[0]; L.y = c0[1] + nn[1] * W.sgn[1]; L.z = c0[2] + nn[2] * W.sgn[2];
      L.tmin = fmaxf(fmaxf(tn_[0], tn_[1]), tn_[2]);
      L.tmax = fminf(fminf(tf_[0], tf_[1]), tf_[2]);
    }
    W.resync = false;
    if (m < r_exit) {                                          // the walker moves on to the first hit that was not latched
      int nn[3];
      float tn_[3], tf_[3];
      times(slot[m], nn, tn_, tf_);
#pragma unroll
      for (int a = 0; a < 3; a++) {
        W.cell[a] = c0[a] + nn[a] * W.sgn[a];
        W.pa[a] = (float)(E[a] + nn[a] * W.sgn[a]) * S.inv_gs_f;
        W.tn[a] = tn_[a];
        W.tf[a] = tf_[a];
      }
      more = true;
    } else more = false;                                       // the ray left the lattice behind the last hit latched
    return m;
  }
  // latch up to `want` (<= 32) hits, lane k taking the k-th; returns the number latched
  __device__ __forceinline__ int gather(LaneSample &L, int want)
  {
#if LRF_WRAY_PARALLEL
    {
#if LRF_WRAY_CHECK
      const typename RayWalker<SPARSE>::State W0 = W.save();
      const bool more0 = more;
      LaneSample Lp = L;
      const int mp = gather_parallel(Lp, want);
      if (mp >= 0) {
        const typename RayWalker<SPARSE>::State W1 = W.save();
        const bool more1 = more;
        W.restore(W0); more = more0;
        int ms = 0;
        while (ms < want && more) {
          if (lane == ms) {
            L.x = W.cell[0]; L.y = W.cell[1]; L.z = W.cell[2];
            L.tmin = fmaxf(fmaxf(W.tn[0], W.tn[1]), W.tn[2]);
            L.tmax = fminf(fminf(W.tf[0], W.tf[1]), W.tf[2]);
          }
          ms++;
          more = W.next();
        }
        bool bad = ms != mp || more != more1;
        if (lane < ms && (L.x != Lp.x || L.y != Lp.y || L.z != Lp.z || __float_as_uint(L.tmin) != __float_as_uint(Lp.tmin) || __float_as_uint(L.tmax) != __float_as_uint(Lp.tmax))) bad = true;
        if (more && (W.cell[0] != W1.cell[0] || W.cell[1] != W1.cell[1] || W.cell[2] != W1.cell[2] || W.tf[0] != W1.tf[0] || W.tf[1] != W1.tf[1] || W.tf[2] != W1.tf[2] ||
                     fmaxf(fmaxf(W.tn[0], W.tn[1]), W.tn[2]) != fmaxf(fmaxf(W1.tn[0], W1.tn[1]), W1.tn[2]) || W.pa[0] != W1.pa[0] || W.pa[1] != W1.pa[1] || W.pa[2] != W1.pa[2])) bad = true;
        if (bad) printf("WRAY MISMATCH lane %d: serial m %d more %d cell %d %d %d t %a %a | parallel m %d more %d cell %d %d %d t %a %a\n", lane, ms, (int)more, L.x, L.y, L.z,
                        L.tmin, L.tmax, mp, (int)more1, Lp.x, Lp.y, Lp.z, Lp.tmin, Lp.tmax);
        return ms;
      }
#else
      const int mp = gather_parallel(L, want);
      if (mp >= 0) return mp;
#endif
    }
#endif
    int m = 0;
    while (m < want && more) {
      if (lane == m) {
        L.x = W.cell[0]; L.y = W.cell[1]; L.z = W.cell[2];
        L.tmin = fmaxf(fmaxf(W.tn[0], W.tn[1]), W.tn[2]);
        L.tmax = fminf(fminf(W.tf[0], W.tf[1]), W.tf[2]);
      }
      m++;
      more = W.next();
    }
    return m;
  }
  // geometry + density + transmittance of this lane's sample (lanes >= m get tr = 1); carry_t = t of the previous
  // group's last sample (-1 before the first sample of the ray), updated
  __device__ __forceinline__ void shade_density(LaneSample &L, int m, float &carry_t)
  {
    const bool act = lane < m;
    const float t = (L.tmin + L.tmax) * 0.5f;                  // ref :142
    float tp = __shfl_up_sync(0xffffffffu, t, 1);
    if (lane == 0) tp = carry_t;
    if (tp < 0.0f) tp = L.tmin;                                // :139
    float delta = t - tp;                                      // :144
    if (delta > S.step_gate_f)                                 // :145-146
      delta = (float)((double)delta - (double)((int)((double)delta / S.step) - 1) * S.step);
    carry_t = __shfl_sync(0xffffffffu, t, max(m - 1, 0));
    L.delta = delta;
    L.tr = 1.0f;
    L.dens_raw = -1.0f;
    L.sigma = 0.0f;
    if (!act) return;
    const int c[3] = {L.x, L.y, L.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float bmin = (float)c[a] * S.inv_gs_f;
      const float p = W.o[a] + t * W.d[a];                     // :148
      L.lf[a] = (p - bmin) * S.gs_f;                           // :150
    }
    L.base = (uint32_t)L.x + (uint32_t)L.y * S.gs + (uint32_t)L.z * S.gs * S.gs;
    const float *p = S.dens + L.base;
    const uint32_t sy = S.gs, sz = (uint32_t)S.gs * S.gs;
    const float d000 = __ldg(p), d100 = __ldg(p + 1), d010 = __ldg(p + sy), d110 = __ldg(p + sy + 1);
    const float d001 = __ldg(p + sz), d101 = __ldg(p + sz + 1), d011 = __ldg(p + sz + sy), d111 = __ldg(p + sz + sy + 1);
    const float xy00 = lm_lerp(d000, d100, L.lf[0]), xy10 = lm_lerp(d010, d110, L.lf[0]);
    const float xy01 = lm_lerp(d001, d101, L.lf[0]), xy11 = lm_lerp(d011, d111, L.lf[0]);
    const float xyz0 = lm_lerp(xy00, xy10, L.lf[1]), xyz1 = lm_lerp(xy01, xy11, L.lf[1]);
#if LRF_WRAY_PREFETCH
    if (!GREY) {
      // the colour of the 8 corners is wanted a few hundred cycles from now (behind the density lerps, expf and the 32-step throughput
      // product): ask for its lines now — the kernel is bound by DRAM latency (71 % of the stall samples on the load scoreboard, issue
      // slots 19 % busy), the gathers of a lane cannot all be in flight at once at 128 registers
#pragma unroll
      for (int cn = 0; cn < 8; cn++) {
        const float4 *q = S.col + (size_t)(L.base + (cn & 1) + ((cn >> 1) & 1) * sy + ((cn >> 2) & 1) * sz) * kColF4;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(q + 2));
      }
    }
#endif
    L.dens_raw = lm_lerp(xyz0, xyz1, L.lf[2]);                 // :155-167, bit-exact
    L.sigma = L.dens_raw < 0.0f ? 0.0f : L.dens_raw;           // :170-171
    L.tr = expf_libm(-L.sigma * delta, etab);                  // :175
  }
  // the reference's left-to-right throughput product over the group; T_run = throughput entering the group, updated.
  // Returns the number of samples of this group that are shaded (the break at :131 may cut it short; *broke says so).
  __device__ __forceinline__ int throughput(LaneSample &L, int m, float &T_run, bool &broke)
  {
    broke = false;
    int n_valid = m;
    float run = T_run;
    L.T = run;
    for (int j = 0; j < m; j++) {
      const float trj = __shfl_sync(0xffffffffu, L.tr, j);
      if (lane == j) L.T = run;
      run = run * trj;                                         // :189
      if (run < 0.1f) {                                        // :131 — checked before the NEXT sample
        n_valid = j + 1;
        broke = true;
        break;
      }
    }
    T_run = run;
    return n_valid;
  }
  // colour of this lane's sample (lanes < n_valid with a non-zero weight); fills rgb / raw / s / c and the corner weights
  __device__ __forceinline__ void shade_colour(LaneSample &L)
  {
    const float fx = L.lf[0], fy = L.lf[1], fz = L.lf[2];
    const float gx = 1.0f - fx, gy = 1.0f - fy, gz = 1.0f - fz;
    const float w00 = gx * gy, w10 = fx * gy, w01 = gx * fy, w11 = fx * fy;
    L.w[0] = w00 * gz; L.w[1] = w10 * gz; L.w[2] = w01 * gz; L.w[3] = w11 * gz;
    L.w[4] = w00 * fz; L.w[5] = w10 * fz; L.w[6] = w01 * fz; L.w[7] = w11 * fz;
    if (GREY) { L.c[0] = L.c[1] = L.c[2] = 1.0f; L.rgb[0] = L.rgb[1] = L.rgb[2] = 0.0f; L.raw = 0.0f; L.s = 0.0f; return; }
    const uint32_t sy = S.gs, sz = (uint32_t)S.gs * S.gs;
    float R = 0.0f, G = 0.0f, B = 0.0f, Q = 0.0f;
#pragma unroll
    for (int cn = 0; cn < 8; cn++) {
      const uint32_t idx = L.base + (cn & 1) + ((cn >> 1) & 1) * sy + ((cn >> 2) & 1) * sz;
      const float4 *q = S.col + (size_t)idx * kColF4;
      const float4 a = __ldg(q), b = __ldg(q + 1), e = __ldg(q + 2);
      float qq = a.w * Y[0];
      qq = fmaf(b.x, Y[1], qq); qq = fmaf(b.y, Y[2], qq); qq = fmaf(b.z, Y[3], qq); qq = fmaf(b.w, Y[4], qq);
      qq = fmaf(e.x, Y[5], qq); qq = fmaf(e.y, Y[6], qq); qq = fmaf(e.z, Y[7], qq); qq = fmaf(e.w, Y[8], qq);
      R = fmaf(L.w[cn], a.x, R); G = fmaf(L.w[cn], a.y, G); B = fmaf(L.w[cn], a.z, B); Q = fmaf(L.w[cn], qq, Q);
    }
    L.rgb[0] = R; L.rgb[1] = G; L.rgb[2] = B;
    L.raw = Q;
    L.s = lm_clamp(Q, 0.0f, 1.0f);                             // :181
    L.c[0] = lm_clamp(R * L.s, 0.0f, 1.0f); L.c[1] = lm_clamp(G * L.s, 0.0f, 1.0f); L.c[2] = lm_clamp(B * L.s, 0.0f, 1.0f);
  }
};

// the whole backward of one ray by one warp; ref = {r,g,b} target (already grey-mixed by the caller if GREY).
// pass 0 = forward (sample count, final throughput, per-channel colour totals in fp64), pass 1 = the adjoint of every
// sample in forward order:  a_tr_k = -T_k (c_k . aC) + [ (suffix_k . aC) + aT_n T_n ] / tr_k  (last sample: aT_n T_k),
// the closed form of the reverse sweep  aT <- tr aT + (1 - tr)(c . aC);  suffix_k = total - prefix_k.
template <bool SPARSE, bool GREY>
__device__ void warp_ray_backward(const SceneDev &S, const uint64_t *etab, const float *o, const float *d, const float *ref, HitRec *buf,
                                  uint32_t cap, uint32_t *slots = nullptr)
{
  const int lane = threadIdx.x & 31;
  WarpRay<SPARSE, GREY> R(S, etab, slots);
  if (!GREY) sh_eval_2(d, R.Y);
  double ctot[3] = {0.0, 0.0, 0.0};
  float aC[3] = {0.0f, 0.0f, 0.0f}, Tn = 0.0f, aTn = 0.0f;
  uint32_t n_shaded = 0;
  uint32_t gathered = 0;                                        // hits latched by pass 0 (the recorded part of the list)
  for (int pass = 0; pass < 2; pass++) {
    const bool replay = pass == 1 && buf != nullptr;            // pass 1 reads the hit list pass 0 recorded
    if (!replay && !R.begin(o, d)) return;                      // ref :688: empty hit list, no loss term
    float T_run = 1.0f, carry_t = -1.0f;
    double cpre[3] = {0.0, 0.0, 0.0};
    uint32_t left = pass ? n_shaded : 0xffffffffu, count = 0, done = 0;
    while (left) {
      LaneSample L;
      int m;
      if (replay) {
        m = (int)min(left, 32u);
        if (lane < m) {
          const HitRec r = buf[done + lane];
          unpack_cell(r.cell, L.x, L.y, L.z);
          L.tmin = r.tmin; L.tmax = r.tmax;
        }
        done += (uint32_t)m;
      } else {
        m = R.gather(L, (int)min(left, 32u));
        if (pass == 0 && buf != nullptr) {
          if (gathered + (uint32_t)m > cap) buf = nullptr;      // list outgrew the slice: pass 1 walks again
          else if (lane < m) {
            HitRec r;
            r.cell = pack_cell(L.x, L.y, L.z); r.tmin = L.tmin; r.tmax = L.tmax;
            buf[gathered + lane] = r;
          }
          gathered += (uint32_t)m;
        }
      }
      if (m == 0) break;
      R.shade_density(L, m, carry_t);
      bool broke;
      const int nv = R.throughput(L, m, T_run, broke);
      const bool act = lane < nv;
      const bool open = act && !(L.dens_raw < 0.0f);           // relu closed: sigma = 0, tr = 1, weight 0, no adjoint (:170-171)
      const float wgt = open ? L.T * (1.0f - L.tr) : 0.0f;     // :187
      double term[3] = {0.0, 0.0, 0.0};
      // the colour is needed by the adjoint even when wgt == 0: d colour / d tr = -T c does not vanish with (1 - tr)
      if (pass ? open : (wgt != 0.0f)) {
        R.shade_colour(L);
        if (wgt != 0.0f) {
#pragma unroll
          for (int ch = 0; ch < 3; ch++) term[ch] = (double)(wgt * L.c[ch]);
        }
      }
      if (pass == 0) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) ctot[ch] += warp_sum(term[ch]);
        count += (uint32_t)nv;
        if (broke || !R.more) break;
        continue;
      }
      double pre[3];
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        pre[ch] = cpre[ch] + warp_incl_scan(term[ch], lane);   // prefix INCLUDING this sample
        cpre[ch] = __shfl_sync(0xffffffffu, pre[ch], 31);
      }
      const bool last = (left <= 32u) && (lane == m - 1);
      if (open) {
        const float dot = L.c[0] * aC[0] + L.c[1] * aC[1] + L.c[2] * aC[2];
        float a_tr;
        if (last) a_tr = -L.T * dot + aTn * L.T;
        else {
          const double suffix = (ctot[0] - pre[0]) * (double)aC[0] + (ctot[1] - pre[1]) * (double)aC[1] + (ctot[2] - pre[2]) * (double)aC[2];
          a_tr = -L.T * dot + (float)((suffix + (double)aTn * (double)Tn) / (double)L.tr);
        }
        const float sigma = L.sigma;
        // d tr / d sigma = -delta tr;  cauchy: d/dsigma log(1 + 2 sigma^2) = 4 sigma / (1 + 2 sigma^2), weight 0.0005 (:267)
        const float a_dens = -L.delta * L.tr * a_tr + 0.0005f * (4.0f * sigma / (1.0f + 2.0f * sigma * sigma));
        const uint32_t sy = S.gs, sz = (uint32_t)S.gs * S.gs;
        if (a_dens != 0.0f) {
#pragma unroll
          for (int cn = 0; cn < 8; cn++)
            red_add_f32(S.gdens + L.base + (cn & 1) + ((cn >> 1) & 1) * sy + ((cn >> 2) & 1) * sz, L.w[cn] * a_dens);
        }
        if (!GREY && wgt != 0.0f) {
          float a_rgb[3] = {0.0f, 0.0f, 0.0f}, a_s = 0.0f;
#pragma unroll
          for (int ch = 0; ch < 3; ch++) {
            const float a_c = wgt * aC[ch];
            const float u = L.rgb[ch] * L.s;
            if (0.0f < u && u <= 1.0f) { a_rgb[ch] = L.s * a_c; a_s = fmaf(L.rgb[ch], a_c, a_s); }   // clamp passes iff 0 < u <= 1
          }
          const float a_raw = (0.0f < L.raw && L.raw <= 1.0f) ? a_s : 0.0f;
          if (a_raw != 0.0f || a_rgb[0] != 0.0f || a_rgb[1] != 0.0f || a_rgb[2] != 0.0f) {
            float a_sh[9];
#pragma unroll
            for (int j = 0; j < 9; j++) a_sh[j] = R.Y[j] * a_raw;
#pragma unroll
            for (int cn = 0; cn < 8; cn++) {
              float4 *q = S.gcol + (size_t)(L.base + (cn & 1) + ((cn >> 1) & 1) * sy + ((cn >> 2) & 1) * sz) * kColF4;
              const float ww = L.w[cn];
              red_add_v4(q, ww * a_rgb[0], ww * a_rgb[1], ww * a_rgb[2], ww * a_sh[0]);
              red_add_v4(q + 1, ww * a_sh[1], ww * a_sh[2], ww * a_sh[3], ww * a_sh[4]);
              red_add_v4(q + 2, ww * a_sh[5], ww * a_sh[6], ww * a_sh[7], ww * a_sh[8]);
            }
          }
        }
      }
      left -= (uint32_t)m;
    }
    if (pass == 0) {
      n_shaded = count;
      if (n_shaded == 0) return;
      Tn = T_run;
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const float craw = (float)ctot[ch];
        const float cc = lm_clamp(craw, 0.0f, 1.0f);           // :193
        const float diff = cc - ref[ch];                       // :266
        const float sg = diff < 0.0f ? -1.0f : 1.0f;
        aC[ch] = (0.0f < craw && craw <= 1.0f) ? sg : 0.0f;
      }
      // d/dT of 0.0001f * (log(T + 1e-3) + log(1 - T + 1e-3))   (:267, double arithmetic there)
      aTn = (float)((double)0.0001f * (1.0 / ((double)Tn + 1e-3) - 1.0 / ((double)(1.0f - Tn) + 1e-3)));
    }
  }
}

}  // namespace lrf
