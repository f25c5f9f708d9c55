// example_tracer.cpp — bodies of the drop-in tracer class: one C-ABI call each (see example_tracer.h here).
#include "example_tracer.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "Image2d.h"   // LiteImage::SaveBMP — the reference's own writer (ref external/LiteMath/Image2d.cpp:519-533), host side, after the path

namespace {

void must(int rc, const char *what)
{
  if (rc == 0) return;
  std::fprintf(stderr, "RayMarcherExample::%s failed: %s\n", what, lrf_last_error());
  std::abort();
}

size_t cells_of(size_t gs) { return gs * gs * gs; }

}  // namespace

size_t RayMarcherExample::BoxList::size() const { return ctx ? lrf_box_count(ctx) : 0; }

std::vector<int32_t> RayMarcherExample::BoxList::xyz() const
{
  std::vector<int32_t> v(size() * 3);
  if (!v.empty()) must(lrf_get_boxes(ctx, v.data()), "boxes.xyz");
  return v;
}

RayMarcherExample::RayMarcherExample(int device)
{
  must(lrf_create(device, &ctx), "RayMarcherExample (no usable sm_100 device: there is no CPU fallback)");
  boxes.ctx = ctx;
  // same defaults as the reference constructor (ref example_tracer.h:43-49)
  m_worldViewInv = inverse4x4(lookAt(float3(0, 1.5, -3), float3(0, 0, 0), float3(0, 1, 0)));
  m_worldViewProjInv = inverse4x4(perspectiveMatrix(90.0f, 1.0f, 0.1f, 100.0f));
  bb.min = float3(0, 0, 0);
  bb.max = float3(1, 1, 1);
  UpdateMembersPlainData();
}

RayMarcherExample::~RayMarcherExample()
{
  if (ctx) lrf_destroy(ctx);
  ctx = nullptr;
}

void RayMarcherExample::InitGrid(const float _gridSize)
{
  gridSize = (size_t)_gridSize;
  must(lrf_init_grid(ctx, (int)gridSize), "InitGrid");
  grid.clear();      // mirrors are stale now
  grid_d.clear();
}

void RayMarcherExample::InitGrad() { must(lrf_init_grad(ctx), "InitGrad"); }

void RayMarcherExample::SetBoundingBox(const float3 boxMin, const float3 boxMax)
{
  bb.min = boxMin;
  bb.max = boxMax;
  const float mn[3] = {boxMin.x, boxMin.y, boxMin.z}, mx[3] = {boxMax.x, boxMax.y, boxMax.z};
  must(lrf_set_bbox(ctx, mn, mx), "SetBoundingBox");
}

void RayMarcherExample::zeroGrad() { must(lrf_zero_grad(ctx), "zeroGrad"); }

void RayMarcherExample::optimizerInit()
{
  lr = 0.15f; beta_1 = 0.9f; beta_2 = 0.999f; eps = 1e-8;   // ref example_tracer.h:120-123
  must(lrf_set_adam_hparams(ctx, lr, beta_1, beta_2, eps), "optimizerInit");
  must(lrf_adam_init(ctx), "optimizerInit");
  V.clear();
  S.clear();
}

void RayMarcherExample::optimizerStep(int iter) { must(lrf_adam_step(ctx, iter, 0), "optimizerStep"); }
void RayMarcherExample::optimizerStepDensity(int iter) { must(lrf_adam_step(ctx, iter, 1), "optimizerStepDensity"); }
void RayMarcherExample::optimizerStepSH(int iter) { must(lrf_adam_step(ctx, iter, 2), "optimizerStepSH"); }

void RayMarcherExample::UpsampleGrid()
{
  must(lrf_upsample_grid(ctx), "UpsampleGrid");
  gridSize = (size_t)lrf_grid_size(ctx);
  grid.clear();
  grid_d.clear();
}

void RayMarcherExample::UpsampleOctree() { must(lrf_upsample_boxes(ctx), "UpsampleOctree"); }
void RayMarcherExample::RebuildOctree() { must(lrf_rebuild_boxes(ctx), "RebuildOctree"); }

void RayMarcherExample::SetWorldViewMatrix(const float4x4 &a_mat)
{
  m_worldViewInv = inverse4x4(a_mat);
  must(lrf_set_view_inv(ctx, (const float *)&m_worldViewInv), "SetWorldViewMatrix");
}

void RayMarcherExample::SetWorldViewMProjatrix(const float4x4 &a_mat)
{
  m_worldViewProjInv = inverse4x4(a_mat);
  must(lrf_set_proj_inv(ctx, (const float *)&m_worldViewProjInv), "SetWorldViewMProjatrix");
}

void RayMarcherExample::kernel2D_RayMarch(uint32_t *out_color, uint32_t width, uint32_t height, bool greyscale)
{
  must(lrf_render(ctx, out_color, width, height, greyscale ? 1 : 0), "kernel2D_RayMarch");
  float t[4] = {0, 0, 0, 0};
  if (lrf_last_time_ms(ctx, "RayMarch", t) == 0) rayMarchTime = t[0];
}

void RayMarcherExample::kernel2D_RayMarchViews(uint32_t *out_color, const float4x4 *a_viewMats, uint32_t n_views, uint32_t width, uint32_t height, bool greyscale)
{
  std::vector<float4x4> inv(n_views);
  for (uint32_t v = 0; v < n_views; v++) inv[v] = inverse4x4(a_viewMats[v]);   // LiteMath's own inverse, as in SetWorldViewMatrix: bit-identical matrices
  must(lrf_render_views(ctx, n_views, (const float *)inv.data(), out_color, width, height, greyscale ? 1 : 0), "kernel2D_RayMarchViews");
}

void RayMarcherExample::kernel2D_RayMarchGrad(uint32_t width, uint32_t height, float4 *res, bool greyscale)
{
  must(lrf_render_grad(ctx, width, height, (const float *)res, greyscale ? 1 : 0), "kernel2D_RayMarchGrad");
}

void RayMarcherExample::kernel2D_TVGrad() { must(lrf_tv_grad(ctx), "kernel2D_TVGrad"); }

void RayMarcherExample::RayMarch(uint32_t *, uint32_t, uint32_t) {}

void RayMarcherExample::CommitDeviceData()
{
  const size_t n = cells_of(gridSize);
  if (n == 0) return;
  if (grid.size() == n) must(lrf_grid_upload(ctx, (const float *)grid.data()), "CommitDeviceData(grid)");
  if (grid_d.size() == n) must(lrf_grad_upload(ctx, (const float *)grid_d.data()), "CommitDeviceData(grid_d)");
  if (V.size() == n * CellSize) must(lrf_adam_state_upload(ctx, 0, V.data()), "CommitDeviceData(V)");
  if (S.size() == n * CellSize) must(lrf_adam_state_upload(ctx, 1, S.data()), "CommitDeviceData(S)");
}

void RayMarcherExample::UpdateMembersPlainData()
{
  must(lrf_set_view_inv(ctx, (const float *)&m_worldViewInv), "UpdateMembersPlainData");
  must(lrf_set_proj_inv(ctx, (const float *)&m_worldViewProjInv), "UpdateMembersPlainData");
  const float mn[3] = {bb.min.x, bb.min.y, bb.min.z}, mx[3] = {bb.max.x, bb.max.y, bb.max.z};
  must(lrf_set_bbox(ctx, mn, mx), "UpdateMembersPlainData");
  must(lrf_set_adam_hparams(ctx, lr, beta_1, beta_2, eps), "UpdateMembersPlainData");
  must(lrf_sync(ctx), "UpdateMembersPlainData");
}

void RayMarcherExample::GetExecutionTime(const char *a_funcName, float a_out[4])
{
  a_out[0] = a_out[1] = a_out[2] = a_out[3] = 0.0f;
  lrf_last_time_ms(ctx, a_funcName, a_out);   // unknown names leave zeros
}

void RayMarcherExample::syncToHost(bool with_grad, bool with_adam)
{
  const size_t n = cells_of(gridSize);
  if (n == 0) return;
  grid.resize(n);
  must(lrf_grid_download(ctx, (float *)grid.data()), "syncToHost(grid)");
  if (with_grad) {
    grid_d.resize(n);
    must(lrf_grad_download(ctx, (float *)grid_d.data()), "syncToHost(grid_d)");
  }
  if (with_adam) {
    V.resize(n * CellSize);
    S.resize(n * CellSize);
    must(lrf_adam_state_download(ctx, 0, V.data()), "syncToHost(V)");
    must(lrf_adam_state_download(ctx, 1, S.data()), "syncToHost(S)");
  }
}

void RayMarcherExample::SaveModelDat(const char *path) { must(lrf_model_save(ctx, path), "SaveModelDat"); }
void RayMarcherExample::LoadModelDat(const char *path) { must(lrf_model_load(ctx, path), "LoadModelDat"); }

// ref example_tracer.cpp:603-625: backward over the whole image, forward render, BMP of the render.  The reference
// zeroes *loss and never computes it; so does this.
void L1Loss(float *loss, float4 *ref, uint *gen, int width, int height, RayMarcherExample *pImpl, const char *fileName, bool greyscale)
{
  *loss = 0.0f;
  pImpl->kernel2D_RayMarchGrad((uint32_t)width, (uint32_t)height, ref, greyscale);
  pImpl->kernel2D_RayMarch(gen, (uint32_t)width, (uint32_t)height, greyscale);
  if (fileName && fileName[0]) LiteImage::SaveBMP(fileName, gen, width, height);
}
