// lrf_camera.cpp — camera helpers for harnesses that do not have LiteMath (the Python bench / tools): bit-identical restatements of
// the LiteMath routines main.cpp uses (main.cpp:78,99-102; LiteMath.h:1362-1368,1556-1569,1571-1653,1770-1815), same operation
// order, compiled without FMA contraction.  Harness code: NOT part of liblrf_b200.so (the C++ drop-in uses LiteMath itself).
#include <cmath>
#include <cstring>

extern "C" {
static void hm_mul_vec(float *res, const float *B, const float *V)   // LiteMath.h:1362-1368
{
  res[0] = V[0] * B[0] + V[1] * B[4] + V[2] * B[8] + V[3] * B[12];
  res[1] = V[0] * B[1] + V[1] * B[5] + V[2] * B[9] + V[3] * B[13];
  res[2] = V[0] * B[2] + V[1] * B[6] + V[2] * B[10] + V[3] * B[14];
  res[3] = V[0] * B[3] + V[1] * B[7] + V[2] * B[11] + V[3] * B[15];
}
void lrf_mat4_mul(const float a[16], const float b[16], float out[16])   // LiteMath.h:1556-1569
{
  float tmp[16];
  for (int c = 0; c < 4; c++) hm_mul_vec(tmp + 4 * c, a, b + 4 * c);
  std::memcpy(out, tmp, sizeof(tmp));
}
void lrf_cam_perspective(float fovy, float aspect, float zNear, float zFar, float out[16])   // LiteMath.h:1797-1815
{
  const float ymax = zNear * tanf(fovy * 3.14159265358979323846f / 360.0f);
  const float xmax = ymax * aspect;
  const float left = -xmax, right = +xmax, bottom = -ymax, top = +ymax;
  const float temp = 2.0f * zNear, temp2 = right - left, temp3 = top - bottom, temp4 = zFar - zNear;
  const float m[16] = {temp / temp2, 0.0f, 0.0f, 0.0f, 0.0f, temp / temp3, 0.0f, 0.0f,
                       (right + left) / temp2, (top + bottom) / temp3, (-zFar - zNear) / temp4, (float)-1.0,
                       0.0f, 0.0f, (-temp * zFar) / temp4, 0.0f};
  std::memcpy(out, m, sizeof(m));
}
static void hm_normalize3(float *v)
{
  const float len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const float inv = float(1) / len;
  v[0] = v[0] * inv; v[1] = v[1] * inv; v[2] = v[2] * inv;
}
static void hm_cross(const float *a, const float *b, float *r)   // LiteMath.h:829-834
{
  const float t0 = a[0] * b[1] - a[1] * b[0], t1 = a[1] * b[2] - a[2] * b[1], t2 = a[2] * b[0] - a[0] * b[2];
  r[0] = t1; r[1] = t2; r[2] = t0;
}
void lrf_cam_orbit_view(int k, int n_views, float out[16])   // main.cpp:99-102
{
  const float eye[3] = {0.0f, 0.0f, 1.3f}, up[3] = {0.0f, 1.0f, 0.0f};
  float x[3], y[3], z[3] = {eye[0] - 0.0f, eye[1] - 0.0f, eye[2] - 0.0f};   // lookAt LiteMath.h:1770-1795
  hm_normalize3(z);
  hm_cross(up, z, x);
  hm_cross(z, x, y);
  hm_normalize3(x);
  hm_normalize3(y);
  const float la[16] = {x[0], y[0], z[0], 0.0f, x[1], y[1], z[1], 0.0f, x[2], y[2], z[2], 0.0f,
                        -x[0] * eye[0] - x[1] * eye[1] - x[2] * eye[2], -y[0] * eye[0] - y[1] * eye[1] - y[2] * eye[2],
                        -z[0] * eye[0] - z[1] * eye[1] - z[2] * eye[2], 1.0f};
  const float DEG_TO_RAD = float(3.14159265358979323846f) / 180.0f;
  const float phi = -float(360.0 / n_views * k) * DEG_TO_RAD;
  const float rot[16] = {+cosf(phi), 0.0f, -sinf(phi), 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, +sinf(phi), 0.0f, +cosf(phi), 0.0f, 0.0f, 0.0f, 0.0f, 1.0f};
  const float tr[16] = {1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, -0.5f, -0.5f, -0.5f, 1.0f};
  float tmp[16];
  lrf_mat4_mul(la, rot, tmp);
  lrf_mat4_mul(tmp, tr, out);
}
#define HM(m, r, c) (m)[(c) * 4 + (r)]
void lrf_mat4_inverse(const float m1[16], float out[16])   // LiteMath.h:1571-1653 (cofactor expansion, same operation order)
{
  float t[12], m[16];
  t[0] = HM(m1, 2, 2) * HM(m1, 3, 3); t[1] = HM(m1, 2, 3) * HM(m1, 3, 2); t[2] = HM(m1, 2, 1) * HM(m1, 3, 3);
  t[3] = HM(m1, 2, 3) * HM(m1, 3, 1); t[4] = HM(m1, 2, 1) * HM(m1, 3, 2); t[5] = HM(m1, 2, 2) * HM(m1, 3, 1);
  t[6] = HM(m1, 2, 0) * HM(m1, 3, 3); t[7] = HM(m1, 2, 3) * HM(m1, 3, 0); t[8] = HM(m1, 2, 0) * HM(m1, 3, 2);
  t[9] = HM(m1, 2, 2) * HM(m1, 3, 0); t[10] = HM(m1, 2, 0) * HM(m1, 3, 1); t[11] = HM(m1, 2, 1) * HM(m1, 3, 0);
  HM(m, 0, 0) = t[0] * HM(m1, 1, 1) + t[3] * HM(m1, 1, 2) + t[4] * HM(m1, 1, 3);
  HM(m, 0, 0) -= t[1] * HM(m1, 1, 1) + t[2] * HM(m1, 1, 2) + t[5] * HM(m1, 1, 3);
  HM(m, 1, 0) = t[1] * HM(m1, 1, 0) + t[6] * HM(m1, 1, 2) + t[9] * HM(m1, 1, 3);
  HM(m, 1, 0) -= t[0] * HM(m1, 1, 0) + t[7] * HM(m1, 1, 2) + t[8] * HM(m1, 1, 3);
  HM(m, 2, 0) = t[2] * HM(m1, 1, 0) + t[7] * HM(m1, 1, 1) + t[10] * HM(m1, 1, 3);
  HM(m, 2, 0) -= t[3] * HM(m1, 1, 0) + t[6] * HM(m1, 1, 1) + t[11] * HM(m1, 1, 3);
  HM(m, 3, 0) = t[5] * HM(m1, 1, 0) + t[8] * HM(m1, 1, 1) + t[11] * HM(m1, 1, 2);
  HM(m, 3, 0) -= t[4] * HM(m1, 1, 0) + t[9] * HM(m1, 1, 1) + t[10] * HM(m1, 1, 2);
  HM(m, 0, 1) = t[1] * HM(m1, 0, 1) + t[2] * HM(m1, 0, 2) + t[5] * HM(m1, 0, 3);
  HM(m, 0, 1) -= t[0] * HM(m1, 0, 1) + t[3] * HM(m1, 0, 2) + t[4] * HM(m1, 0, 3);
  HM(m, 1, 1) = t[0] * HM(m1, 0, 0) + t[7] * HM(m1, 0, 2) + t[8] * HM(m1, 0, 3);
  HM(m, 1, 1) -= t[1] * HM(m1, 0, 0) + t[6] * HM(m1, 0, 2) + t[9] * HM(m1, 0, 3);
  HM(m, 2, 1) = t[3] * HM(m1, 0, 0) + t[6] * HM(m1, 0, 1) + t[11] * HM(m1, 0, 3);
  HM(m, 2, 1) -= t[2] * HM(m1, 0, 0) + t[7] * HM(m1, 0, 1) + t[10] * HM(m1, 0, 3);
  HM(m, 3, 1) = t[4] * HM(m1, 0, 0) + t[9] * HM(m1, 0, 1) + t[10] * HM(m1, 0, 2);
  HM(m, 3, 1) -= t[5] * HM(m1, 0, 0) + t[8] * HM(m1, 0, 1) + t[11] * HM(m1, 0, 2);
  t[0] = HM(m1, 0, 2) * HM(m1, 1, 3); t[1] = HM(m1, 0, 3) * HM(m1, 1, 2); t[2] = HM(m1, 0, 1) * HM(m1, 1, 3);
  t[3] = HM(m1, 0, 3) * HM(m1, 1, 1); t[4] = HM(m1, 0, 1) * HM(m1, 1, 2); t[5] = HM(m1, 0, 2) * HM(m1, 1, 1);
  t[6] = HM(m1, 0, 0) * HM(m1, 1, 3); t[7] = HM(m1, 0, 3) * HM(m1, 1, 0); t[8] = HM(m1, 0, 0) * HM(m1, 1, 2);
  t[9] = HM(m1, 0, 2) * HM(m1, 1, 0); t[10] = HM(m1, 0, 0) * HM(m1, 1, 1); t[11] = HM(m1, 0, 1) * HM(m1, 1, 0);
  HM(m, 0, 2) = t[0] * HM(m1, 3, 1) + t[3] * HM(m1, 3, 2) + t[4] * HM(m1, 3, 3);
  HM(m, 0, 2) -= t[1] * HM(m1, 3, 1) + t[2] * HM(m1, 3, 2) + t[5] * HM(m1, 3, 3);
  HM(m, 1, 2) = t[1] * HM(m1, 3, 0) + t[6] * HM(m1, 3, 2) + t[9] * HM(m1, 3, 3);
  HM(m, 1, 2) -= t[0] * HM(m1, 3, 0) + t[7] * HM(m1, 3, 2) + t[8] * HM(m1, 3, 3);
  HM(m, 2, 2) = t[2] * HM(m1, 3, 0) + t[7] * HM(m1, 3, 1) + t[10] * HM(m1, 3, 3);
  HM(m, 2, 2) -= t[3] * HM(m1, 3, 0) + t[6] * HM(m1, 3, 1) + t[11] * HM(m1, 3, 3);
  HM(m, 3, 2) = t[5] * HM(m1, 3, 0) + t[8] * HM(m1, 3, 1) + t[11] * HM(m1, 3, 2);
  HM(m, 3, 2) -= t[4] * HM(m1, 3, 0) + t[9] * HM(m1, 3, 1) + t[10] * HM(m1, 3, 2);
  HM(m, 0, 3) = t[2] * HM(m1, 2, 2) + t[5] * HM(m1, 2, 3) + t[1] * HM(m1, 2, 1);
  HM(m, 0, 3) -= t[4] * HM(m1, 2, 3) + t[0] * HM(m1, 2, 1) + t[3] * HM(m1, 2, 2);
  HM(m, 1, 3) = t[8] * HM(m1, 2, 3) + t[0] * HM(m1, 2, 0) + t[7] * HM(m1, 2, 2);
  HM(m, 1, 3) -= t[6] * HM(m1, 2, 2) + t[9] * HM(m1, 2, 3) + t[1] * HM(m1, 2, 0);
  HM(m, 2, 3) = t[6] * HM(m1, 2, 1) + t[11] * HM(m1, 2, 3) + t[3] * HM(m1, 2, 0);
  HM(m, 2, 3) -= t[10] * HM(m1, 2, 3) + t[2] * HM(m1, 2, 0) + t[7] * HM(m1, 2, 1);
  HM(m, 3, 3) = t[10] * HM(m1, 2, 2) + t[4] * HM(m1, 2, 0) + t[9] * HM(m1, 2, 1);
  HM(m, 3, 3) -= t[8] * HM(m1, 2, 1) + t[11] * HM(m1, 2, 2) + t[5] * HM(m1, 2, 0);
  const float k = 1.0f / (HM(m1, 0, 0) * HM(m, 0, 0) + HM(m1, 0, 1) * HM(m, 1, 0) + HM(m1, 0, 2) * HM(m, 2, 0) + HM(m1, 0, 3) * HM(m, 3, 0));
  for (int i = 0; i < 16; i++) out[i] = m[i] * k;
}

}  // extern "C"
