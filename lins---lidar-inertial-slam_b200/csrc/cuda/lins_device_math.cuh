// Device-side small math for the IESKF kernels (sm_100a).  f64 geometry in the operation order of the
// reference's Eigen expressions (see DESIGN.md "precision contract"); this TU is compiled with -fmad=false so
// no multiply-add is contracted: the f32 distance expression and the f64->f32 stores round exactly like the
// reference's x86-64 SSE2 build (lins/CMakeLists.txt:3, -O3 without -march=native => no FMA).
// Reference functions mirrored: lins/include/math_utils.h axis2Quat :43-73, Quat2axis :75-88, skew :197-204,
// Rinvleft :304-321; Eigen QuaternionBase::_transformVector / toRotationMatrix.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace lins_dev {

struct d3 { double x, y, z; };
struct q4 { double w, x, y, z; };
struct m3 { double m[9]; };  // row-major

__device__ __forceinline__ d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 add3(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 sub3(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 scl3(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ d3 div3(d3 a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }
// Eigen 3-element redux order x0 + (x1 + x2)
__device__ __forceinline__ double dot3(d3 a, d3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
__device__ __forceinline__ double norm3(d3 a) { return sqrt(dot3(a, a)); }
__device__ __forceinline__ d3 cross3(d3 a, d3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ q4 qmul(q4 a, q4 b) {
  q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
__device__ __forceinline__ q4 qnormalized(q4 q) {
  double n = sqrt((q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w));
  q4 r; r.w = q.w / n; r.x = q.x / n; r.y = q.y / n; r.z = q.z / n;
  return r;
}
__device__ __forceinline__ q4 qinverse(q4 q) {
  double n2 = (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);
  q4 r;
  if (n2 > 0) { r.w = q.w / n2; r.x = -q.x / n2; r.y = -q.y / n2; r.z = -q.z / n2; }
  else { r.w = r.x = r.y = r.z = 0; }
  return r;
}
// QuaternionBase::_transformVector: uv = 2*(q.vec x v); v + w*uv + q.vec x uv
__device__ __forceinline__ d3 qrot(q4 q, d3 v) {
  d3 qv = mk3(q.x, q.y, q.z);
  d3 uv = cross3(qv, v);
  uv = add3(uv, uv);
  d3 c = cross3(qv, uv);
  return mk3((v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z);
}
__device__ __forceinline__ m3 qtoR(q4 q) {
  m3 r;
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r.m[0] = 1.0 - (tyy + tzz); r.m[1] = txy - twz;         r.m[2] = txz + twy;
  r.m[3] = txy + twz;         r.m[4] = 1.0 - (txx + tzz); r.m[5] = tyz - twx;
  r.m[6] = txz - twy;         r.m[7] = tyz + twx;         r.m[8] = 1.0 - (txx + tyy);
  return r;
}
__device__ __forceinline__ double wrap_pi(double x) {
  const double PI = 3.14159265358979323846;
  while (x >= PI) x -= 2.0 * PI;
  while (x < -PI) x += 2.0 * PI;
  return x;
}
// math_utils.h:61-73 (+ :43-59)
__device__ __forceinline__ q4 axis2Quat(d3 vec) {
  q4 q;
  double theta = norm3(vec);
  if (theta < 1e-10) { q.w = 1.0; q.x = q.y = q.z = 0.0; return q; }
  d3 t = div3(vec, theta);
  double sn, cs;
  sincos(theta / 2.0, &sn, &cs);
  q.w = cs; q.x = t.x * sn; q.y = t.y * sn; q.z = t.z * sn;
  return q;
}
// math_utils.h:75-88
__device__ __forceinline__ d3 Quat2axis(q4 q) {
  double mag = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  d3 v = mk3(q.x, q.y, q.z);
  if (mag >= 1e-10) {
    v = div3(v, mag);
    v = scl3(wrap_pi(2.0 * atan2(mag, q.w)), v);
  }
  return v;
}
// math_utils.h:304-321
__device__ __forceinline__ m3 Rinvleft(d3 axis) {
  m3 r;
  double theta = norm3(axis);
  if (theta < 1e-10) {
    for (int i = 0; i < 9; ++i) r.m[i] = 0.0;
    r.m[0] = r.m[4] = r.m[8] = 1.0;
    return r;
  }
  double h = theta / 2.0;
  d3 a = div3(axis, theta);
  double sn, cs;
  sincos(h, &sn, &cs);
  double s = h * (cs / sn);
  double oms = 1.0 - s;
  d3 sa = scl3(oms, a);
  // s*I + ((1-s)*a)*a^T - h*skew(a)
  r.m[0] = (s + sa.x * a.x);           r.m[1] = (sa.x * a.y) - h * (-a.z);  r.m[2] = (sa.x * a.z) - h * (a.y);
  r.m[3] = (sa.y * a.x) - h * (a.z);   r.m[4] = (s + sa.y * a.y);           r.m[5] = (sa.y * a.z) - h * (-a.x);
  r.m[6] = (sa.z * a.x) - h * (-a.y);  r.m[7] = (sa.z * a.y) - h * (a.x);   r.m[8] = (s + sa.z * a.z);
  return r;
}

// exact f32 squared distance, flann L2_Simple order ((dx*dx)+dy*dy)+dz*dz, never contracted
__device__ __forceinline__ float sqdist_f32(float qx, float qy, float qz, float tx, float ty, float tz) {
  float dx = __fsub_rn(qx, tx), dy = __fsub_rn(qy, ty), dz = __fsub_rn(qz, tz);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  r = __fadd_rn(r, __fmul_rn(dz, dz));
  return r;
}

// ---- 1-D TMA (cp.async.bulk) + mbarrier helpers ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bytes must be a multiple of 16, src/dst 16-B aligned
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

}  // namespace lins_dev
