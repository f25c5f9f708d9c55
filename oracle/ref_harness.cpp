// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// Thin extern "C" harness around the UNMODIFIED reference tracer.  This translation unit
// textually includes /root/reference/example_tracer/example_tracer.cpp (the include path
// is supplied by oracle/Makefile; nothing is copied into this repo), so that every
// function of the reference — including the file-static EyeRayDir / transform_ray3f /
// RealColorToUint32 — is callable from ctypes.  It provides the one symbol the Enzyme
// clang plugin would normally provide (__enzyme_autodiff), as a stub that aborts: the
// reference's backward cannot be built here (no clang-17 / Enzyme), see DESIGN.md.
//
// Built twice by oracle/Makefile into oracle/_ref/:
//   libref_O3.so : -O3 -ffp-contract=off -fopenmp   (fair CPU baseline, forward golden vectors)
//   libref_O0.so : -O0 -fopenmp                     (flags as shipped, CMakeLists.txt:62; the only
//                                                   build whose TVRegularisation is meaningful)
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <omp.h>
#include <sstream>
#include <iostream>

#include "example_tracer/example_tracer.cpp"  // the reference, where it lies
// stb_image exactly as main.cpp instantiates it (main.cpp:12-13: a C++ TU, so stb's `pow(float, float)` is the float overload)
#define STB_IMAGE_IMPLEMENTATION
#include <stb_image.h>
#include "Image2d.h"

float __enzyme_autodiff(...)
{
  std::fprintf(stderr, "ref_harness: __enzyme_autodiff is not available (no Enzyme plugin)\n");
  std::abort();
}

namespace {
struct Silence  // the reference prints progress lines to std::cout; keep test logs clean
{
  std::streambuf *old;
  std::ostringstream sink;
  Silence() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Silence() { std::cout.rdbuf(old); }
};
float4x4 from16(const float *m)  // m = 16 floats, column-major (m_col[0..3])
{
  float4x4 r;
  std::memcpy((void *)&r, m, 16 * sizeof(float));
  return r;
}
void to16(const float4x4 &a, float *m) { std::memcpy(m, (const void *)&a, 16 * sizeof(float)); }
}  // namespace

extern "C" {

void *ref_create() { return new RayMarcherExample(); }
void ref_destroy(void *h) { delete (RayMarcherExample *)h; }

// InitGrid + the rest of main.cpp:49-58
void ref_init_grid(void *h, int gs)
{
  Silence s;
  auto *p = (RayMarcherExample *)h;
  p->InitGrid((float)gs);
  p->optimizerInit();
  p->InitGrad();
  p->SetBoundingBox(float3(0, 0, 0), float3(1, 1, 1));
}
// grid only, no boxes/octree (for Adam/TV/upsample-grid tests at sizes where the octree build is slow)
void ref_init_grid_nobox(void *h, int gs)
{
  auto *p = (RayMarcherExample *)h;
  p->gridSize = gs;
  p->grid.assign((size_t)gs * gs * gs, Cell{});
  p->optimizerInit();
  p->InitGrad();
  p->SetBoundingBox(float3(0, 0, 0), float3(1, 1, 1));
}
void ref_set_bbox(void *h, const float *mn, const float *mx)
{
  ((RayMarcherExample *)h)->SetBoundingBox(float3(mn[0], mn[1], mn[2]), float3(mx[0], mx[1], mx[2]));
}
size_t ref_grid_size(void *h) { return ((RayMarcherExample *)h)->gridSize; }
float *ref_grid_ptr(void *h) { return (float *)((RayMarcherExample *)h)->grid.data(); }
float *ref_grad_ptr(void *h) { return (float *)((RayMarcherExample *)h)->grid_d.data(); }
float *ref_adam_v_ptr(void *h) { return ((RayMarcherExample *)h)->V.data(); }
float *ref_adam_s_ptr(void *h) { return ((RayMarcherExample *)h)->S.data(); }
void ref_get_adam_hparams(void *h, float *out4)
{
  auto *p = (RayMarcherExample *)h;
  out4[0] = p->lr; out4[1] = p->beta_1; out4[2] = p->beta_2; out4[3] = p->eps;
}
size_t ref_box_count(void *h) { return ((RayMarcherExample *)h)->boxes.size(); }
void ref_get_boxes(void *h, double *out6)  // Min[3], Max[3] per box
{
  auto *p = (RayMarcherExample *)h;
  for (size_t i = 0; i < p->boxes.size(); i++)
    for (int a = 0; a < 3; a++) {
      out6[i * 6 + a] = p->boxes[i].Min[a];
      out6[i * 6 + 3 + a] = p->boxes[i].Max[a];
    }
}
void ref_set_boxes(void *h, const double *in6, size_t n)
{
  auto *p = (RayMarcherExample *)h;
  p->boxes.resize(n);
  for (size_t i = 0; i < n; i++)
    p->boxes[i] = OrthoTree::BoundingBox3D{{in6[i * 6], in6[i * 6 + 1], in6[i * 6 + 2]},
                                           {in6[i * 6 + 3], in6[i * 6 + 4], in6[i * 6 + 5]}};
  p->octree = OrthoTree::OctreeBoxC::Create(p->boxes);
}

// ---- camera: exactly the caller-side formulas of main.cpp:78,99-104 ----
void ref_make_perspective(float fovy, float aspect, float zn, float zf, float *out16)
{
  to16(perspectiveMatrix(fovy, aspect, zn, zf), out16);
}
void ref_make_orbit_view(int k, int n_views, float *out16)  // main.cpp:99-102 with IMAGE_LOOP = n_views
{
  float4x4 viewMat = lookAt(float3(0.0, 0.0, 1.3), float3(0.0, 0.0, 0.0), float3(0.0, 1.0, 0.0)) *
                     rotate4x4Y(-float(360.0 / n_views * k) * DEG_TO_RAD) *
                     translate4x4(float3(-0.5, -0.5, -0.5));
  to16(viewMat, out16);
}
void ref_inverse4x4(const float *in16, float *out16) { to16(inverse4x4(from16(in16)), out16); }
void ref_set_view_matrix(void *h, const float *m16) { ((RayMarcherExample *)h)->SetWorldViewMatrix(from16(m16)); }
void ref_set_proj_matrix(void *h, const float *m16) { ((RayMarcherExample *)h)->SetWorldViewMProjatrix(from16(m16)); }
void ref_get_inv_matrices(void *h, float *view_inv16, float *proj_inv16)
{
  auto *p = (RayMarcherExample *)h;
  to16(p->m_worldViewInv, view_inv16);
  to16(p->m_worldViewProjInv, proj_inv16);
}

// ---- the path itself ----
void ref_render(void *h, uint32_t *out, uint32_t w, uint32_t hgt, int grey)
{
  ((RayMarcherExample *)h)->kernel2D_RayMarch(out, w, hgt, grey != 0);
}

// ray generation of kernel2D_RayMarch lines 635-639
void ref_gen_ray(void *h, uint32_t x, uint32_t y, uint32_t w, uint32_t hgt, float *o, float *d)
{
  auto *p = (RayMarcherExample *)h;
  float3 rayDir = EyeRayDir((float(x) + 0.5f) / float(w), (float(y) + 0.5f) / float(hgt), p->m_worldViewProjInv);
  float3 rayPos = float3(0.0f, 0.0f, 0.0f);
  transform_ray3f(p->m_worldViewInv, &rayPos, &rayDir);
  for (int a = 0; a < 3; a++) { o[a] = rayPos[a]; d[a] = rayDir[a]; }
}

// octree.RayIntersectedAll exactly as called at example_tracer.cpp:641-643; returns the count,
// writes at most cap ids
size_t ref_ray_hits(void *h, const float *o, const float *d, uint64_t *ids, size_t cap)
{
  auto *p = (RayMarcherExample *)h;
  std::vector<size_t> hits = p->octree.RayIntersectedAll(OrthoTree::Point3D{o[0], o[1], o[2]},
                                                          OrthoTree::Point3D{d[0], d[1], d[2]}, 0.0);
  for (size_t i = 0; i < hits.size() && i < cap; i++) ids[i] = hits[i];
  return hits.size();
}

static void shade_prefix(RayMarcherExample *p, const float *o, const float *d, const std::vector<size_t> &ids,
                         int grey, float *rgb, float *T, float *cauchy)
{
  float step = (p->bb.max[0] - p->bb.min[0]) / (p->gridSize);  // example_tracer.cpp:649
  float3 c = RayGridIntersection(float3(o[0], o[1], o[2]), float3(d[0], d[1], d[2]), step, ids, p->boxes,
                                 p->grid.data(), p->gridSize, *T, *cauchy, grey != 0);
  rgb[0] = c[0]; rgb[1] = c[1]; rgb[2] = c[2];
}

// RayGridIntersection on an explicit id list
void ref_shade(void *h, const float *o, const float *d, const uint64_t *ids, size_t n, int grey, float *rgb,
               float *T, float *cauchy)
{
  std::vector<size_t> v(ids, ids + n);
  shade_prefix((RayMarcherExample *)h, o, d, v, grey, rgb, T, cauchy);
}

// number of samples RayGridIntersection shades before its `throughput < 0.1` break
// (example_tracer.cpp:131): smallest m with T(prefix m) < 0.1, else n.  Uses only the reference's
// own function (bisection over the prefix length; T(prefix) is non-increasing).
static size_t n_shaded_of(RayMarcherExample *p, const float *o, const float *d, const std::vector<size_t> &hits, int grey)
{
  float rgb[3], T, cauchy;
  shade_prefix(p, o, d, hits, grey, rgb, &T, &cauchy);
  if (!(T < 0.1)) return hits.size();
  size_t lo = 0, hi = hits.size();  // T(lo) >= 0.1 (T(0) = 1), T(hi) < 0.1
  while (hi - lo > 1) {
    size_t mid = (lo + hi) / 2;
    std::vector<size_t> pre(hits.begin(), hits.begin() + mid);
    shade_prefix(p, o, d, pre, grey, rgb, &T, &cauchy);
    if (T < 0.1) hi = mid; else lo = mid;
  }
  return hi;
}

// Whole-image debug taps, one entry per pixel (row-major, y*w+x):
//   n_hit, n_shaded, rgbtc = {r,g,b,T,cauchy} floats, first `keep` box ids per ray (-1 padded; keep may be 0)
void ref_render_debug(void *h, uint32_t w, uint32_t hgt, int grey, uint32_t *n_hit, uint32_t *n_shaded,
                      float *rgbtc, int64_t *ids_out, uint32_t keep)
{
  auto *p = (RayMarcherExample *)h;
#pragma omp parallel for schedule(dynamic, 1)
  for (uint32_t y = 0; y < hgt; y++)
    for (uint32_t x = 0; x < w; x++) {
      size_t pix = (size_t)y * w + x;
      float o[3], d[3];
      ref_gen_ray(h, x, y, w, hgt, o, d);
      std::vector<size_t> hits = p->octree.RayIntersectedAll(OrthoTree::Point3D{o[0], o[1], o[2]},
                                                              OrthoTree::Point3D{d[0], d[1], d[2]}, 0.0);
      n_hit[pix] = (uint32_t)hits.size();
      float *out = rgbtc + pix * 5;
      out[0] = out[1] = out[2] = out[3] = out[4] = 0.0f;
      n_shaded[pix] = 0;
      if (hits.size()) {
        shade_prefix(p, o, d, hits, grey, out, out + 3, out + 4);
        n_shaded[pix] = (uint32_t)n_shaded_of(p, o, d, hits, grey);
      }
      for (uint32_t i = 0; i < keep; i++) ids_out[pix * keep + i] = i < hits.size() ? (int64_t)hits[i] : -1;
    }
}

// RayGridLoss (example_tracer.cpp:258-268) on an explicit id list, with step as kernel2D_RayMarchGrad:692
float ref_ray_loss(void *h, const float *ref_rgba, const float *o, const float *d, const uint64_t *ids, size_t n, int grey)
{
  auto *p = (RayMarcherExample *)h;
  std::vector<size_t> v(ids, ids + n);
  float4 ref(ref_rgba[0], ref_rgba[1], ref_rgba[2], ref_rgba[3]);
  float3 ro(o[0], o[1], o[2]), rd(d[0], d[1], d[2]);
  double step = (p->bb.max[0] - p->bb.min[0]) / (p->gridSize);
  Cell *g = p->grid.data();
  size_t gs = p->gridSize;
  float fw = 0, fh = 0;
  bool gr = grey != 0;
  return RayGridLoss(ref, ro, rd, step, v, p->boxes, g, gs, fw, fh, gr);
}

// TVRegularisation (example_tracer.cpp:196-248).  Only meaningful from the -O0 build (UB at :241-245).
float ref_tv(void *h, size_t x, size_t y, size_t z)
{
  auto *p = (RayMarcherExample *)h;
  Cell *g = p->grid.data();
  size_t gs = p->gridSize;
  return TVRegularisation(x, y, z, g, gs);
}

void ref_zero_grad(void *h) { ((RayMarcherExample *)h)->zeroGrad(); }
void ref_adam_step(void *h, int iter, int which)  // 0 all, 1 density, 2 SH (example_tracer.h:128-206)
{
  auto *p = (RayMarcherExample *)h;
  if (which == 0) p->optimizerStep(iter);
  else if (which == 1) p->optimizerStepDensity(iter);
  else p->optimizerStepSH(iter);
}
void ref_upsample_grid(void *h) { Silence s; ((RayMarcherExample *)h)->UpsampleGrid(); ((RayMarcherExample *)h)->InitGrad(); }
void ref_upsample_octree(void *h) { Silence s; ((RayMarcherExample *)h)->UpsampleOctree(); }
void ref_rebuild_octree(void *h) { Silence s; ((RayMarcherExample *)h)->RebuildOctree(); }

uint32_t ref_pack_color(const float *rgba) { return RealColorToUint32(float4(rgba[0], rgba[1], rgba[2], rgba[3])); }
float ref_expf(float x) { return exp(x); }   // whatever `exp(float)` resolves to inside the reference TU (:175)
float ref_logf(float x) { return log(x); }   // idem for `log(float)` (:173)
int ref_omp_threads() { return omp_get_max_threads(); }
void ref_set_omp_threads(int n) { omp_set_num_threads(n); }   // torchrun exports OMP_NUM_THREADS=1: the CPU arm asks for the cores back
// main.cpp:72,96: stbi_set_flip_vertically_on_load(true); stbi_loadf(path, &w, &h, &c, 4).  out (may be NULL) receives w*h*4 floats.
int ref_stbi_loadf(const char *path, int flip, float *out, size_t cap_floats, int *w, int *hgt)
{
  stbi_set_flip_vertically_on_load(flip);
  int c = 4;
  float *p = stbi_loadf(path, w, hgt, &c, 4);
  if (!p) return 1;
  const size_t n = (size_t)(*w) * (*hgt) * 4;
  if (out && n <= cap_floats) std::memcpy(out, p, n * sizeof(float));
  free(p);
  return 0;
}
// example_tracer.cpp:614
int ref_save_bmp(const char *path, const uint32_t *pixels, int w, int hgt) { return LiteImage::SaveBMP(path, pixels, w, hgt) ? 0 : 1; }

}  // extern "C"
