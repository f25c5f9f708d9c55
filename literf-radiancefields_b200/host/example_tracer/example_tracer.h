// example_tracer.h — link-time drop-in for the reference's tracer class, backed by liblrf_b200.so.
//
// Put this directory AHEAD of the reference's own example_tracer/ on the include path and the reference's
// main.cpp compiles unmodified (`#include "example_tracer/example_tracer.h"`, ref main.cpp:9): same class
// name, same public method signatures (ref example_tracer/example_tracer.h:40-252), same public data-member
// names.  Every method forwards to one C-ABI call of include/lrf.h; nothing here computes on the CPU.
// LiteMath (float3 / float4 / float4x4, inverse4x4, lookAt, ...) is the reference's own header, used as-is
// on the host so that the matrices handed to the device are bit-identical to the reference's.
//
// Differences a maintainer should know about (all forced by "device memory is the source of truth"):
//   * `grid`, `grid_d`, `V`, `S` are host MIRRORS.  They are empty until syncToHost() fills them from the
//     device; after editing them (e.g. the model.dat load of ref main.cpp:54-56) call CommitDeviceData(),
//     which uploads every mirror whose size matches the current grid.
//   * `boxes` is a proxy exposing size() (all ref main.cpp:136 needs); lattice coordinates of the boxes are
//     available through boxes.xyz().
//   * `octree` does not exist: the exact hit list is produced by a voxel DDA on the device.
//   * GetExecutionTime returns CUDA-event milliseconds (the reference returns an uninitialised member).
//   * failures abort() with the library's error text (the reference has no error convention at all).
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "LiteMath.h"
using namespace LiteMath;

#include "lrf.h"

const size_t SH_WIDTH = 9;

struct Cell   // ref example_tracer/example_tracer.h:25-30 — 13 floats, the interchange layout of lrf_grid_upload/download
{
  float density;
  float RGB[3];
  float sh[SH_WIDTH];
};
const size_t CellSize = sizeof(Cell) / sizeof(Cell::density);

struct BoundingBox
{
  float3 min;
  float3 max;
};

class RayMarcherExample
{
public:
  // stand-in for `std::vector<OrthoTree::BoundingBox3D> boxes` (ref example_tracer.h:240)
  class BoxList
  {
  public:
    size_t size() const;
    bool empty() const { return size() == 0; }
    std::vector<int32_t> xyz() const;   // n x 3 lattice coordinates in id order
  private:
    friend class RayMarcherExample;
    lrf_ctx *ctx = nullptr;
  };

  explicit RayMarcherExample(int device = 0);
  ~RayMarcherExample();
  RayMarcherExample(const RayMarcherExample &) = delete;
  RayMarcherExample &operator=(const RayMarcherExample &) = delete;

  void InitGrid(const float _gridSize);                       // ref :51-91
  void InitGrad();                                            // ref :93-96
  void SetBoundingBox(const float3 boxMin, const float3 boxMax);   // ref :98-102
  void zeroGrad();                                            // ref :104-109
  void optimizerInit();                                       // ref :111-126
  void optimizerStep(int iter);                               // ref :128-163
  void optimizerStepDensity(int iter);                        // ref :165-183
  void optimizerStepSH(int iter);                             // ref :185-206
  void UpsampleGrid();                                        // ref example_tracer.cpp:281-323
  void UpsampleOctree();                                      // ref example_tracer.cpp:325-447
  void RebuildOctree();                                       // ref example_tracer.cpp:449-489

  void SetWorldViewMatrix(const float4x4 &a_mat);             // ref :214
  void SetWorldViewMProjatrix(const float4x4 &a_mat);         // ref :215 [sic]

  void kernel2D_RayMarch(uint32_t *out_color, uint32_t width, uint32_t height, bool greyscale = false);       // ref example_tracer.cpp:627-663
  void kernel2D_RayMarchGrad(uint32_t width, uint32_t height, float4 *res, bool greyscale = false);           // ref example_tracer.cpp:665-735
  void kernel2D_TVGrad();                                     // ref example_tracer.cpp:737-749
  void RayMarch(uint32_t *out_color, uint32_t width, uint32_t height);   // ref example_tracer.cpp:751-757 (empty body there too)

  void CommitDeviceData();                                    // ref :223  host mirrors -> device
  void UpdateMembersPlainData();                              // ref :224  POD members (matrices, bb, lr/beta/eps) -> device
  void GetExecutionTime(const char *a_funcName, float a_out[4]);   // ref example_tracer.cpp:759-763

  // device -> host mirrors (no reference counterpart)
  void syncToHost(bool with_grad = true, bool with_adam = true);
  // the raw model.dat dump / load that main.cpp keeps commented out (main.cpp:54-56,139-141): gridSize^3 x Cell, no header
  // not in the reference: n_views frames in one call (frame v at out_color + v * width * height), view matrices as SetWorldViewMatrix takes them;
  // one kernel launch for all views (lrf_render_views).  The current view matrix is left as it is.
  void kernel2D_RayMarchViews(uint32_t *out_color, const float4x4 *a_viewMats, uint32_t n_views, uint32_t width, uint32_t height, bool greyscale = false);
  void SaveModelDat(const char *path);
  void LoadModelDat(const char *path);

  float4x4 m_worldViewProjInv;
  float4x4 m_worldViewInv;
  float rayMarchTime = 0.0f;

  std::vector<Cell> grid;
  std::vector<Cell> grid_d;
  size_t gridSize = 0;

  BoxList boxes;

  float lr = 0.15f, beta_1 = 0.9f, beta_2 = 0.999f, eps = 1e-8f;
  std::vector<float> V;
  std::vector<float> S;

  BoundingBox bb;

  lrf_ctx *ctx = nullptr;   // the C-ABI handle, for callers that want the asynchronous / device-pointer entry points
};

void L1Loss(float *loss, float4 *ref, uint *gen, int width, int height, RayMarcherExample *pImpl, const char *fileName,
            bool greyscale = false);   // ref example_tracer.cpp:603-625
