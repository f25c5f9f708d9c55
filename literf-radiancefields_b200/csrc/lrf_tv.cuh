// lrf_tv.cuh — TV gradient as a z-sweep with TMA-fed density planes (round 2).
//
// ref: TVRegularisation example_tracer.cpp:196-248, kernel2D_TVGrad :737-749 (Enzyme adjoint :250-256; sqrt'(0) := 0).
//
// tv_fused_kernel (round 1) works on 8^3 output tiles: a 12^3 density halo per tile (3.4 loads per node), five stages each over a
// cube whose linear index has to be decomposed per element, 10 stage evaluations per output node — ncu: issue slots 87 % busy, a
// fifth of the instructions integer index arithmetic, 0.36 ms for 201 MB of algorithmic traffic at 256^3.
//
// Here a block owns a 32 x 8 column of nodes and SWEEPS it along z.  Per z step ONE density plane (the 36 x 12 halo rectangle)
// arrives by cp.async.bulk.tensor (3-D tensor map over the density plane array, box {40, 12, 1} — 36 wide plus 4 so that it starts 16-byte aligned; a ring of 4 planes keeps three
// in flight) and the stages run on planes:
//     D(z) -> V(z-1) -> tv cells E(z-2) -> G(z-2) -> node gradient (z-2)
// with two-deep rolling buffers for V, Ez and G.  Every thread keeps the SAME (x, y) items for the whole sweep, so all index
// arithmetic is done once; the halo overhead is (36 * 12) / (32 * 8) = 1.7 loads per node (+ 4 planes per z chunk) and 5 stage
// evaluations per output node.  Boundary: TMA zero-fills outside the grid; the reference clamps the upper neighbour
// (min(a + 1, gs - 1), ref :203-230), so tiles on the upper x / y faces copy the last row / column into the halo and the z
// coordinate of a plane request is clamped; cells below 0 are masked out by every stage that could see them (as in round 1).
#pragma once
#include "lrf_staged.cuh"

namespace lrf {

constexpr int kSwX = 32, kSwY = 8;                       // output tile
constexpr int kSwDX = kSwX + 8, kSwDY = kSwY + 4;        // density halo plane 40 x 12: x from q0x - 4 (the start of a TMA box must be 16-byte aligned
                                                         // in memory: 4 floats), y from q0y - 2; the stages use x >= q0x - 2, i.e. local x + 2
constexpr int kSwVX = kSwX + 3, kSwVY = kSwY + 3;        // V plane 35 x 11             <-> q0 - 2 + l
constexpr int kSwEX = kSwX + 2, kSwEY = kSwY + 2;        // tv cells 34 x 10            <-> q0 - 2 + l
constexpr int kSwGX = kSwX + 1, kSwGY = kSwY + 1;        // G plane 33 x 9              <-> q0 - 1 + l
constexpr int kSwRing = 6;                              // density planes: 2 being read + 4 in flight
constexpr int kSwSlotF = 512;                            // floats per ring slot: 40 * 12 = 480 rounded up to a 128-byte multiple
constexpr int kSwZChunk = 64;                            // z planes per block

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(tm), "r"(bar),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// grid: one block of 256 threads per work item (tile x z chunk), or a persistent grid walking them; nodes with z in [z_begin, z_end)
// get their TV gradient ADDED to gdens with red.global.add (fire and forget: no load to wait for; other kernels may be scattering
// into the same plane at the same time)
//
// Software pipeline over z: in step j the density plane zz = zlo + j lands and FOUR stages run side by side, each on buffers filled
// in earlier steps, so ONE block barrier per step is enough (the staged version above needed four; ncu showed barrier stalls and the
// read-modify-write of gdens as the two largest stall reasons):
//     V(zz-1)  <- D(zz-1), D(zz)                     ring of 4
//     E(zz-3)  <- V(zz-3), V(zz-2)                   Ex, Ey: 2 buffers, Ez: ring of 3
//     G(zz-4)  <- E(zz-4), Ez(zz-5), V(zz-4)         ring of 3
//     out(zz-5) <- G(zz-5), G(zz-6)
__global__ void __launch_bounds__(256) tv_sweep_kernel(const __grid_constant__ CUtensorMap tm, int gs, float *__restrict__ gdens, int z_begin, int z_end,
                                                       int atomic, uint32_t *diag)
{
  (void)atomic;
  __shared__ __align__(128) float sD[kSwRing][kSwSlotF];
  __shared__ float sV[4][kSwVX * kSwVY], sEx[2][kSwEX * kSwEY], sEy[2][kSwEX * kSwEY], sEz[3][kSwEX * kSwEY], sG[3][kSwGX * kSwGY];
  __shared__ __align__(8) uint64_t sBar[kSwRing];
  const int t = threadIdx.x, m = gs - 1;
  const int tiles_x = gs / kSwX, tiles_y = gs / kSwY;
  const uint32_t a_bar = smem_u32(sBar), a_d = smem_u32(&sD[0][0]);
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < kSwRing; s++) mbar_init(a_bar + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int n_work = tiles_x * tiles_y * ((z_end - z_begin + kSwZChunk - 1) / kSwZChunk);
  int jj = 0;                                             // planes this block has requested so far (ring slot and barrier phase run on)
  // ---- the items of this thread within a tile, the same for every tile and every z step ----
  // V: 385 items, E: 340, G: 297 (two per thread for the low thread ids), node gradient: exactly one per thread
  const int v_i[2] = {t, t + 256}, e_i[2] = {t, t + 256}, g_i[2] = {t, t + 256};
  int v_d[2], e_v[2], g_e[2], g_v[2], e_x[2], e_y[2], g_x[2], g_y[2];
  bool v_on[2], e_on[2], g_on[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    v_on[k] = v_i[k] < kSwVX * kSwVY;
    v_d[k] = (v_i[k] % kSwVX + 2) + (v_i[k] / kSwVX) * kSwDX;
    e_on[k] = e_i[k] < kSwEX * kSwEY;
    e_x[k] = e_i[k] % kSwEX; e_y[k] = e_i[k] / kSwEX;
    e_v[k] = e_x[k] + e_y[k] * kSwVX;
    g_on[k] = g_i[k] < kSwGX * kSwGY;
    g_x[k] = g_i[k] % kSwGX; g_y[k] = g_i[k] / kSwGX;
    g_e[k] = (g_x[k] + 1) + (g_y[k] + 1) * kSwEX;
    g_v[k] = (g_x[k] + 1) + (g_y[k] + 1) * kSwVX;
  }
  const int ox = t % kSwX, oy = t / kSwX;
  const int o_g = (ox + 1) + (oy + 1) * kSwGX;
  const size_t plane = (size_t)gs * gs;
  // a block walks the work items blockIdx.x, blockIdx.x + gridDim.x, ...: one item per block when the kernel has the GPU to itself,
  // a few resident blocks per SM when it runs beside the HBM-bound Adam pass
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
    const int bx = work % tiles_x, by = (work / tiles_x) % tiles_y, bc = work / (tiles_x * tiles_y);
    const int q0x = bx * kSwX, q0y = by * kSwY;
    const int zb = z_begin + bc * kSwZChunk, ze = min(zb + kSwZChunk, z_end);
    const int zlo = zb - 2, P = ze - zb + 4;              // density planes zlo .. ze + 1
    const int j0 = jj;
    jj += P;
    auto issue = [&](int j) {                             // thread 0 only
      const int slot = (j0 + j) % kSwRing;
      const uint32_t bar = a_bar + 8 * slot;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot was read (and, on the upper faces, patched) through the generic proxy
      mbar_expect_tx(bar, kSwDX * kSwDY * 4);
      tma_load_3d(a_d + (uint32_t)slot * kSwSlotF * 4, &tm, bar, q0x - 4, q0y - 2, min(max(zlo + j, 0), m));
    };
    bool e_in[2], g_in[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const int a = q0x - 2 + e_x[k], b = q0y - 2 + e_y[k];
      e_in[k] = a >= 0 && b >= 0 && a < m && b < m;        // the tv cell exists laterally (ref loop bounds :739-741)
      g_in[k] = q0x - 1 + g_x[k] >= 0 && q0y - 1 + g_y[k] >= 0;
    }
    const int qx = q0x + ox, qy = q0y + oy;
    const float wx0 = qx == m ? 1.0f : 0.5f, wx1 = qx > 0 ? 0.5f : 0.0f;    // V cell q: weight 0.5, or 1 at the last node (far == near); V cell q - 1: absent at q == 0
    const float wy0 = qy == m ? 1.0f : 0.5f, wy1 = qy > 0 ? 0.5f : 0.0f;
    float *const o_dst = gdens + (size_t)qx + (size_t)qy * gs;
    const bool edge = q0x - 4 + kSwDX - 1 > m || q0y - 2 + kSwDY - 1 > m;
    __syncthreads();                                      // the previous item's last reads of the ring and the stage buffers are done
    if (t == 0)
      for (int j = 0; j < min(kSwRing - 2, P); j++) issue(j);

    for (int j = 0; j < P + 3; j++) {
      const int zz = zlo + j;                             // density plane that lands in this step (j < P)
      if (j < P && t == 0) mbar_wait_one(a_bar + 8 * ((j0 + j) % kSwRing), (uint32_t)((j0 + j) / kSwRing) & 1u, diag, 6u, (uint32_t)j, blockIdx.x);
      __syncthreads();                                    // D(zz) is visible; every buffer written in step j - 1 is complete
      // planes j - 2 and older are consumed (V(zz-1) of step j - 1 was their last reader): refill the ring two steps behind the reads
      if (t == 0 && j + kSwRing - 2 < P) issue(j + kSwRing - 2);
      float *const D1 = sD[(j0 + j) % kSwRing];
      if (j < P && edge) {
        // upper faces: the reference's far neighbour is min(a + 1, gs - 1) — copy the last column / row into the zero-filled halo
        for (int i = t; i < kSwDX * kSwDY; i += 256) {
          const int lx = i % kSwDX, ly = i / kSwDX, gx = q0x - 4 + lx, gy = q0y - 2 + ly;
          if (gx > m || gy > m) D1[i] = D1[(min(gx, m) - (q0x - 4)) + (min(gy, m) - (q0y - 2)) * kSwDX];
        }
        __syncthreads();
      }
      if (j >= 1 && j < P) {
        // V(zz - 1): trilinear at 0.5 over {a, a + 1}^3 in the reference's nesting (getVoxel :203-230)
        const float *const D0 = sD[(j0 + j - 1) % kSwRing];
        float *const V = sV[(j - 1) & 3];
#pragma unroll
        for (int k = 0; k < 2; k++)
          if (v_on[k]) {
            const float *p = D0 + v_d[k], *pz = D1 + v_d[k];
            const float xy00 = lm_lerp(p[0], p[1], 0.5f), xy10 = lm_lerp(p[kSwDX], p[kSwDX + 1], 0.5f);
            const float xy01 = lm_lerp(pz[0], pz[1], 0.5f), xy11 = lm_lerp(pz[kSwDX], pz[kSwDX + 1], 0.5f);
            V[v_i[k]] = lm_lerp(lm_lerp(xy00, xy10, 0.5f), lm_lerp(xy01, xy11, 0.5f), 0.5f);
          }
      }
      if (j >= 3 && j < P + 1) {
        // tv cells E(c), c = zz - 3: lam * (dx, dy, dz) / sqrt(dx^2 + dy^2 + dz^2); 0 where the cell does not exist or the root's argument is 0
        const int c = zz - 3;
        const float *const V0 = sV[(j - 3) & 3], *const V1 = sV[(j - 2) & 3];
        float *const Ex = sEx[(j - 3) & 1], *const Ey = sEy[(j - 3) & 1], *const Ez = sEz[(j - 3) % 3];
        const bool zin = c >= 0 && c < m;
#pragma unroll
        for (int k = 0; k < 2; k++)
          if (e_on[k]) {
            float ex = 0.0f, ey = 0.0f, ez = 0.0f;
            if (zin && e_in[k]) {
              const float *p = V0 + e_v[k];
              const float v = fmaxf(p[0], 0.0f);
              const float dx = v - fmaxf(p[1], 0.0f), dy = v - fmaxf(p[kSwVX], 0.0f), dz = v - fmaxf(V1[e_v[k]], 0.0f);
              const float sum = dx * dx + dy * dy + dz * dz;
              if (sum > 0.0f) {
                const float kk = 0.01f * rsqrtf(sum);
                ex = dx * kk; ey = dy * kk; ez = dz * kk;
              }
            }
            Ex[e_i[k]] = ex; Ey[e_i[k]] = ey; Ez[e_i[k]] = ez;
          }
      }
      if (j >= 5 && j < P + 2) {
        // G(p) = d(sum tv) / dV(p), p = zz - 4: (Ex + Ey + Ez)(p) - Ex(p - ex) - Ey(p - ey) - Ez(p - ez), gated by the relu (:226-227)
        const float *const V0 = sV[(j - 4) & 3], *const Ex = sEx[(j - 4) & 1], *const Ey = sEy[(j - 4) & 1];
        const float *const Ez0 = sEz[(j - 4) % 3], *const Ezm = sEz[(j - 5) % 3];
        float *const G = sG[(j - 4) % 3];
        const bool zin = zz - 4 >= 0;
#pragma unroll
        for (int k = 0; k < 2; k++)
          if (g_on[k]) {
            const int e = g_e[k];
            float g = Ex[e] + Ey[e] + Ez0[e] - Ex[e - 1] - Ey[e - kSwEX] - Ezm[e];
            if (V0[g_v[k]] < 0.0f || !zin || !g_in[k]) g = 0.0f;
            G[g_i[k]] = g;
          }
      }
      const int qz = zz - 5;
      if (j >= 7 && qz >= zb && qz < ze) {
        // node gradient: each node is a corner of up to 8 V cells
        const float *const p = sG[(j - 5) % 3] + o_g, *const pm = sG[(j - 6) % 3] + o_g;
        const float wz0 = qz == m ? 1.0f : 0.5f, wz1 = qz > 0 ? 0.5f : 0.0f;
        float acc = 0.0f;
        acc += wz1 * (wy1 * (wx1 * pm[-kSwGX - 1] + wx0 * pm[-kSwGX]) + wy0 * (wx1 * pm[-1] + wx0 * pm[0]));
        acc += wz0 * (wy1 * (wx1 * p[-kSwGX - 1] + wx0 * p[-kSwGX]) + wy0 * (wx1 * p[-1] + wx0 * p[0]));
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(o_dst + (size_t)qz * plane), "f"(acc) : "memory");
      }
    }
  }
}

}  // namespace lrf
