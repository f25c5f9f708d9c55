/* TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the LiteRF ray-march hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (literf-radiancefields_b200/) never
 * includes, links or calls anything declared here.
 *
 * Parity status:
 *   forward (ray gen, hit set + order, shading, packing, upsample, Adam, TV value):
 *       PINNED — checked bit-for-bit against the reference's own code compiled from
 *       /root/reference (oracle/_ref, tests/test_oracle_vs_ref.py) and against the committed
 *       golden vectors generated from it (tests/golden/), plus the attcs/Octree
 *       RayIntersectedAll known-answer vectors (external/Octree/unittests/general.tests.cpp:1008-1063).
 *   backward (ray-loss adjoint, TV gradient): PARITY UNPINNED against the reference — the
 *       reference generates it with the Enzyme clang plugin (unpinned version, not in the tree,
 *       not buildable here).  The adjoint is hand-derived from example_tracer.cpp:120-268 and
 *       self-certified against central finite differences of an fp64 restatement of the
 *       reference's own loss, which in turn is cross-checked against the compiled RayGridLoss.
 *
 * Conventions: grid = gs^3 nodes x 13 floats, AoS, x fastest (example_tracer.h:25-32,
 * example_tracer.cpp:62-65).  Boxes = unit voxels of the (gs-1)^3 lattice given by integer
 * coordinates; the position of a box in the list is its id (the octree's EntityID, used as the
 * tie-break key, external/Octree/octree.h:856-860).  Matrices = 16 floats column-major.
 */
#ifndef LRF_ORACLE_H
#define LRF_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* camera helpers — main.cpp:78,99-104 + LiteMath.h:1521-1529,1571-1653,1770-1815 */
void orc_perspective(float fovy, float aspect, float zn, float zf, float *out16);
void orc_orbit_view(int k, int n_views, float *out16);
void orc_inverse4x4(const float *in16, float *out16);

void *orc_create(int gs);                 /* InitGrid contents + dense box list + Adam init */
void orc_destroy(void *h);
int orc_grid_size(void *h);
float *orc_grid_ptr(void *h);
float *orc_grad_ptr(void *h);
float *orc_adam_v_ptr(void *h);
float *orc_adam_s_ptr(void *h);
size_t orc_box_count(void *h);
void orc_get_boxes_xyz(void *h, int32_t *xyz);
void orc_set_boxes_xyz(void *h, const int32_t *xyz, size_t n);
void orc_set_bbox(void *h, const float *mn, const float *mx);
void orc_set_view_inv(void *h, const float *m16);
void orc_set_proj_inv(void *h, const float *m16);

void orc_gen_ray(void *h, uint32_t x, uint32_t y, uint32_t w, uint32_t hgt, float *o, float *d);
size_t orc_ray_hits(void *h, const float *o, const float *d, uint64_t *ids, size_t cap);
/* arbitrary boxes, arbitrary dimension (2 or 3): IsRayHit + sort, for the Octree KATs */
size_t orc_ray_hits_generic(int dim, const double *boxes_minmax, size_t n, const double *origin,
                            const double *heading, uint64_t *ids, size_t cap);
size_t orc_shade(void *h, const float *o, const float *d, const uint64_t *ids, size_t n, int grey,
                 float *rgb, float *T, float *cauchy);          /* returns n_shaded */
void orc_render(void *h, uint32_t *out, uint32_t w, uint32_t hgt, int grey);
void orc_render_debug(void *h, uint32_t w, uint32_t hgt, int grey, uint32_t *n_hit, uint32_t *n_shaded,
                      float *rgbtc, int64_t *ids_out, uint32_t keep);
/* packed rays: o,d = n x 3 floats */
void orc_trace_rays(void *h, size_t n, const float *o, const float *d, int grey, uint32_t *n_hit,
                    uint32_t *n_shaded, float *rgbtc, int32_t *vox_xyz_out, uint32_t keep);
/* distinct grid nodes read by at least one shaded sample of an image (SURVEY 8d N_touched) */
uint64_t orc_count_touched_nodes(void *h, uint32_t w, uint32_t hgt, int grey);

float orc_ray_loss(void *h, const float *ref_rgba, const float *o, const float *d, const uint64_t *ids,
                   size_t n, int grey);
double orc_ray_loss_f64(void *h, const double *grid64, const float *ref_rgba, const float *o,
                        const float *d, const uint64_t *ids, size_t n, int grey);
/* adjoint of the ray loss wrt the grid, accumulated (+=) serially */
void orc_ray_grad(void *h, const float *ref_rgba, const float *o, const float *d, const uint64_t *ids,
                  size_t n, int grey);
void orc_ray_grad_f64(void *h, const double *grid64, double *grad64, const float *ref_rgba, const float *o,
                      const float *d, const uint64_t *ids, size_t n, int grey);
void orc_render_grad(void *h, uint32_t w, uint32_t hgt, const float *ref_rgba, int grey);
void orc_rays_grad(void *h, size_t n, const float *o, const float *d, const float *ref_rgba, int grey);

float orc_tv(void *h, size_t x, size_t y, size_t z);
double orc_tv_total_f64(void *h, const double *density64);
void orc_tv_grad(void *h);
void orc_zero_grad(void *h);
void orc_adam_init(void *h);
void orc_adam_step(void *h, int iter, int which);       /* 0 all, 1 density, 2 SH/RGB */
void orc_upsample_grid(void *h);
void orc_upsample_boxes(void *h);
void orc_rebuild_boxes(void *h);

uint32_t orc_pack_color(const float *rgba);
float orc_expf_model(float x, int use_fma);            /* glibc 2.27+ expf algorithm, restated */
void orc_lcg_fill_grid(void *h, float density_scale, int blob);  /* SURVEY 8(d) synthetic grid */
int orc_omp_threads(void);
void orc_set_omp_threads(int n);
/* stbi_loadf's flip + ldr->hdr on decoded RGBA8 rows (ref external/stb_image.h:1241-1258,1856-1876; main.cpp:72,96) */
void orc_ldr_to_hdr_rgba8(const uint8_t *rgba8, uint32_t w, uint32_t hgt, int flip, float *out);
/* LiteImage::SaveBMP file image (ref external/LiteMath/Image2d.cpp:488-533); returns 54 + 3*w*h */
size_t orc_bmp_bytes(const uint32_t *pixels, int w, int hgt, uint8_t *out);
#ifdef __cplusplus
}
#endif
#endif
