// oracle/lins_math.hpp — TEST INFRASTRUCTURE (CPU oracle), not product code.
//
// Dependency-free restatement of the small fixed-size linear algebra the reference's hot path takes from
// Eigen and from lins/include/math_utils.h.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
//
// Parity status: UNPINNED at the third-party boundary.  Eigen is not vendored in /root/reference and is not
// installed here, so operation ORDER inside Eigen expressions (which only matters in the last ulp of f64) is
// restated from Eigen 3.2.92/3.3 as documented below, not checked against a built reference.
//   * 3-element reductions (dot, squaredNorm, 1x3*3x3 product coefficients) use Eigen's non-vectorised
//     unrolled redux order  x0 + (x1 + x2)   (Eigen/src/Core/Redux.h redux_novec_unroller, Vector3d has no
//     packet access because sizeof == 24 is not a multiple of 16).
//   * Quaterniond * Vector3d follows QuaternionBase::_transformVector:  uv = 2*(q.vec x v);
//     v + w*uv + q.vec x uv.
//   * toRotationMatrix follows QuaternionBase::toRotationMatrix (tx=2x ... formulation).
//   * quaternion product is the generic (non-SSE) formula of Eigen's quat_product.
#ifndef LINS_ORACLE_MATH_HPP_
#define LINS_ORACLE_MATH_HPP_

#include <cmath>
#include <cstring>

namespace lins_oracle {

struct V3 {
  double x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  double operator()(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(const V3& a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(double s, const V3& a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator*(const V3& a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(const V3& a, double s) { return V3(a.x / s, a.y / s, a.z / s); }
// Eigen redux order for 3 elements: x0 + (x1 + x2)
inline double dot(const V3& a, const V3& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline double squaredNorm(const V3& a) { return a.x * a.x + (a.y * a.y + a.z * a.z); }
inline double norm(const V3& a) { return std::sqrt(squaredNorm(a)); }
// Eigen MatrixBase::cross
inline V3 cross(const V3& a, const V3& b) {
  return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

struct M3 {
  double m[3][3];  // m[row][col]
  M3() { std::memset(m, 0, sizeof(m)); }
  static M3 Identity() {
    M3 r;
    r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0;
    return r;
  }
  double operator()(int r, int c) const { return m[r][c]; }
  double& operator()(int r, int c) { return m[r][c]; }
};
inline M3 operator*(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + (a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j]);
  return r;
}
inline V3 operator*(const M3& a, const V3& v) {
  return V3(a.m[0][0] * v.x + (a.m[0][1] * v.y + a.m[0][2] * v.z), a.m[1][0] * v.x + (a.m[1][1] * v.y + a.m[1][2] * v.z),
            a.m[2][0] * v.x + (a.m[2][1] * v.y + a.m[2][2] * v.z));
}
// row-vector * matrix : (v^T A)_j
inline V3 rowTimes(const V3& v, const M3& a) {
  return V3(v.x * a.m[0][0] + (v.y * a.m[1][0] + v.z * a.m[2][0]), v.x * a.m[0][1] + (v.y * a.m[1][1] + v.z * a.m[2][1]),
            v.x * a.m[0][2] + (v.y * a.m[1][2] + v.z * a.m[2][2]));
}
inline M3 operator*(double s, const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = s * a.m[i][j];
  return r;
}
inline M3 operator+(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r;
}
inline M3 operator-(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j];
  return r;
}
inline M3 operator-(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = -a.m[i][j];
  return r;
}
inline M3 transpose(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
inline M3 outer(const V3& a, const V3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a(i) * b(j);
  return r;
}

// math_utils.h:197-204  skew
inline M3 skew(const V3& q) {
  M3 a;
  a.m[0][0] = 0;     a.m[0][1] = -q.z;  a.m[0][2] = q.y;
  a.m[1][0] = q.z;   a.m[1][1] = 0;     a.m[1][2] = -q.x;
  a.m[2][0] = -q.y;  a.m[2][1] = q.x;   a.m[2][2] = 0;
  return a;
}

struct Q4 {
  double w, x, y, z;
  Q4() : w(1), x(0), y(0), z(0) {}
  Q4(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}
  V3 vec() const { return V3(x, y, z); }
};
// Eigen generic quat_product
inline Q4 operator*(const Q4& a, const Q4& b) {
  return Q4(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x);
}
inline double squaredNorm(const Q4& q) { return (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w); }
inline Q4 normalized(const Q4& q) {
  double n = std::sqrt(squaredNorm(q));
  return Q4(q.w / n, q.x / n, q.y / n, q.z / n);
}
// QuaternionBase::inverse: conjugate / squaredNorm
inline Q4 inverse(const Q4& q) {
  double n2 = squaredNorm(q);
  if (n2 > 0) return Q4(q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2);
  return Q4(0, 0, 0, 0);
}
// QuaternionBase::_transformVector
inline V3 operator*(const Q4& q, const V3& v) {
  V3 uv = cross(q.vec(), v);
  uv = uv + uv;
  V3 c = cross(q.vec(), uv);
  return V3((v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z);
}
// QuaternionBase::toRotationMatrix
inline M3 toRotationMatrix(const Q4& q) {
  M3 r;
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r.m[0][0] = 1.0 - (tyy + tzz);  r.m[0][1] = txy - twz;          r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz;          r.m[1][1] = 1.0 - (txx + tzz);  r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy;          r.m[2][1] = tyz + twx;          r.m[2][2] = 1.0 - (txx + tyy);
  return r;
}

// math_utils.h:27-37  wrap_pi
inline double wrap_pi(double x) {
  while (x >= double(M_PI)) x -= double(2.0 * M_PI);
  while (x < double(-M_PI)) x += double(2.0 * M_PI);
  return x;
}

// math_utils.h:43-59  axis2Quat(axis, theta).  (The theta<1e-10 branch falls through in the reference.)
inline Q4 axis2Quat(const V3& axis, double theta) {
  Q4 q;
  double magnitude = std::sin(theta / 2.0f);
  q.w = std::cos(theta / 2.0f);
  q.x = axis.x * magnitude;
  q.y = axis.y * magnitude;
  q.z = axis.z * magnitude;
  return q;
}
// math_utils.h:61-73  axis2Quat(vec)
inline Q4 axis2Quat(const V3& vec) {
  double theta = norm(vec);
  if (theta < 1e-10) return Q4(1.0, 0, 0, 0);
  V3 tmp = vec / theta;
  return axis2Quat(tmp, theta);
}
// math_utils.h:75-88  Quat2axis
inline V3 Quat2axis(const Q4& q) {
  double axis_magnitude = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  V3 vec(q.x, q.y, q.z);
  if (axis_magnitude >= 1e-10) {
    vec = vec / axis_magnitude;
    vec = vec * wrap_pi(2.0 * std::atan2(axis_magnitude, q.w));
  }
  return vec;
}
// math_utils.h:304-321  Rinvleft
inline M3 Rinvleft(const V3& axis) {
  double theta = norm(axis);
  if (theta < 1e-10) return M3::Identity();
  double half_theta = theta / 2.0;
  V3 a = axis / norm(axis);
  double cot_half_theta = std::cos(half_theta) / std::sin(half_theta);
  double s = half_theta * cot_half_theta;
  // s*I + (1-s)*a*a^T - half_theta*skew(a)   (evaluated left to right, element-wise)
  M3 ans = (s * M3::Identity() + outer((1.0 - s) * a, a)) - half_theta * skew(a);
  return ans;
}
// math_utils.h:131-149  rpy2Quat  (the trailing Q.normalized() discards its result)
inline Q4 rpy2Quat(const V3& rpy) {
  double halfYaw = rpy.z * 0.5, halfPitch = rpy.y * 0.5, halfRoll = rpy.x * 0.5;
  double cosYaw = std::cos(halfYaw), sinYaw = std::sin(halfYaw);
  double cosPitch = std::cos(halfPitch), sinPitch = std::sin(halfPitch);
  double cosRoll = std::cos(halfRoll), sinRoll = std::sin(halfRoll);
  Q4 Q;
  Q.x = sinRoll * cosPitch * cosYaw - cosRoll * sinPitch * sinYaw;
  Q.y = cosRoll * sinPitch * cosYaw + sinRoll * cosPitch * sinYaw;
  Q.z = cosRoll * cosPitch * sinYaw - sinRoll * sinPitch * cosYaw;
  Q.w = cosRoll * cosPitch * cosYaw + sinRoll * sinPitch * sinYaw;
  return Q;
}
// math_utils.h:190-196 R2rpy / Q2rpy
inline V3 R2rpy(const M3& R) {
  V3 rpy;
  rpy.y = std::atan2(-R(2, 0), std::sqrt(R(2, 1) * R(2, 1) + R(2, 2) * R(2, 2)));
  rpy.x = std::atan2(R(2, 1) / std::cos(rpy.y), R(2, 2) / std::cos(rpy.y));
  rpy.z = std::atan2(R(1, 0) / std::cos(rpy.y), R(0, 0) / std::cos(rpy.y));
  return rpy;
}

}  // namespace lins_oracle
#endif
