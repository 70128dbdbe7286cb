// Host-side small math for the CPU stages that stay on the CPU (IMU propagation, feature extraction, the
// StateEstimator shim).  Mirrors the API names of the reference's lins/include/math_utils.h
// (axis2Quat :43-73, Quat2axis :75-88, skew :197-204, Rinvleft :304-321, rpy2Quat :131-149, R2rpy :190-196)
// on dependency-free types — Eigen is not available in this image and must not leak through the C-ABI.
// PRODUCT code: never includes anything from oracle/.
#ifndef LINS_HOST_MATH_UTILS_HPP_
#define LINS_HOST_MATH_UTILS_HPP_

#include <array>
#include <cmath>

namespace lins {

struct V3D {
  std::array<double, 3> d{{0, 0, 0}};
  V3D() = default;
  V3D(double x, double y, double z) : d{{x, y, z}} {}
  double x() const { return d[0]; }
  double y() const { return d[1]; }
  double z() const { return d[2]; }
  double& operator()(int i) { return d[i]; }
  double operator()(int i) const { return d[i]; }
  double squaredNorm() const { return d[0] * d[0] + d[1] * d[1] + d[2] * d[2]; }
  double norm() const { return std::sqrt(squaredNorm()); }
  void setZero() { d = {{0, 0, 0}}; }
  V3D& operator+=(const V3D& o) { for (int i = 0; i < 3; ++i) d[i] += o.d[i]; return *this; }
};
inline V3D operator+(V3D a, const V3D& b) { for (int i = 0; i < 3; ++i) a.d[i] += b.d[i]; return a; }
inline V3D operator-(V3D a, const V3D& b) { for (int i = 0; i < 3; ++i) a.d[i] -= b.d[i]; return a; }
inline V3D operator-(V3D a) { for (int i = 0; i < 3; ++i) a.d[i] = -a.d[i]; return a; }
inline V3D operator*(double s, V3D a) { for (int i = 0; i < 3; ++i) a.d[i] *= s; return a; }
inline V3D operator*(V3D a, double s) { return s * a; }
inline V3D operator/(V3D a, double s) { for (int i = 0; i < 3; ++i) a.d[i] /= s; return a; }
inline double dot(const V3D& a, const V3D& b) { return a.d[0] * b.d[0] + a.d[1] * b.d[1] + a.d[2] * b.d[2]; }
inline V3D cross(const V3D& a, const V3D& b) {
  return V3D(a.d[1] * b.d[2] - a.d[2] * b.d[1], a.d[2] * b.d[0] - a.d[0] * b.d[2], a.d[0] * b.d[1] - a.d[1] * b.d[0]);
}

struct M3D {
  std::array<double, 9> a{{0, 0, 0, 0, 0, 0, 0, 0, 0}};  // row-major
  double& operator()(int r, int c) { return a[3 * r + c]; }
  double operator()(int r, int c) const { return a[3 * r + c]; }
  static M3D Identity() { M3D m; m(0, 0) = m(1, 1) = m(2, 2) = 1.0; return m; }
  M3D transpose() const { M3D t; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t(r, c) = (*this)(c, r); return t; }
};
inline M3D operator*(const M3D& x, const M3D& y) {
  M3D r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += x(i, k) * y(k, j); r(i, j) = s; }
  return r;
}
inline V3D operator*(const M3D& m, const V3D& v) {
  V3D r;
  for (int i = 0; i < 3; ++i) r(i) = m(i, 0) * v(0) + m(i, 1) * v(1) + m(i, 2) * v(2);
  return r;
}
inline M3D operator*(double s, M3D m) { for (auto& e : m.a) e *= s; return m; }
inline M3D operator+(M3D x, const M3D& y) { for (int i = 0; i < 9; ++i) x.a[i] += y.a[i]; return x; }
inline M3D operator-(M3D x, const M3D& y) { for (int i = 0; i < 9; ++i) x.a[i] -= y.a[i]; return x; }
inline M3D operator-(M3D x) { for (auto& e : x.a) e = -e; return x; }

// quaternion, Eigen coefficient order (x, y, z, w) so it memcpy's to/from the C-ABI state layout
struct Q4D {
  std::array<double, 4> c{{0, 0, 0, 1}};
  Q4D() = default;
  Q4D(double w, double x, double y, double z) : c{{x, y, z, w}} {}  // Eigen ctor argument order (w,x,y,z)
  double x() const { return c[0]; }
  double y() const { return c[1]; }
  double z() const { return c[2]; }
  double w() const { return c[3]; }
  double& x() { return c[0]; }
  double& y() { return c[1]; }
  double& z() { return c[2]; }
  double& w() { return c[3]; }
  void setIdentity() { c = {{0, 0, 0, 1}}; }
  V3D vec() const { return V3D(c[0], c[1], c[2]); }
  double squaredNorm() const { return c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]; }
  Q4D normalized() const { double n = std::sqrt(squaredNorm()); return Q4D(c[3] / n, c[0] / n, c[1] / n, c[2] / n); }
  Q4D inverse() const { double n = squaredNorm(); return Q4D(c[3] / n, -c[0] / n, -c[1] / n, -c[2] / n); }
  M3D toRotationMatrix() const {
    const double x2 = c[0] + c[0], y2 = c[1] + c[1], z2 = c[2] + c[2];
    const double wx = x2 * c[3], wy = y2 * c[3], wz = z2 * c[3];
    const double xx = x2 * c[0], xy = y2 * c[0], xz = z2 * c[0], yy = y2 * c[1], yz = z2 * c[1], zz = z2 * c[2];
    M3D r;
    r(0, 0) = 1 - (yy + zz); r(0, 1) = xy - wz;       r(0, 2) = xz + wy;
    r(1, 0) = xy + wz;       r(1, 1) = 1 - (xx + zz); r(1, 2) = yz - wx;
    r(2, 0) = xz - wy;       r(2, 1) = yz + wx;       r(2, 2) = 1 - (xx + yy);
    return r;
  }
};
inline Q4D operator*(const Q4D& a, const Q4D& b) {
  return Q4D(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
             a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
             a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
             a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
}
inline V3D operator*(const Q4D& q, const V3D& v) {
  V3D t = 2.0 * cross(q.vec(), v);
  return v + q.w() * t + cross(q.vec(), t);
}

namespace math_utils {

inline int sign(double x) { return x >= 0 ? 1 : -1; }
inline double wrap_pi(double x) {
  while (x >= M_PI) x -= 2.0 * M_PI;
  while (x < -M_PI) x += 2.0 * M_PI;
  return x;
}
inline M3D skew(const V3D& q) {
  M3D a;
  a(0, 1) = -q(2); a(0, 2) = q(1);
  a(1, 0) = q(2);  a(1, 2) = -q(0);
  a(2, 0) = -q(1); a(2, 1) = q(0);
  return a;
}
inline Q4D axis2Quat(const V3D& vec) {
  double theta = vec.norm();
  if (theta < 1e-10) return Q4D();
  V3D ax = vec / theta;
  double m = std::sin(theta / 2.0);
  return Q4D(std::cos(theta / 2.0), ax(0) * m, ax(1) * m, ax(2) * m);
}
inline V3D Quat2axis(const Q4D& q) {
  double mag = std::sqrt(q.x() * q.x() + q.y() * q.y() + q.z() * q.z());
  V3D v(q.x(), q.y(), q.z());
  if (mag >= 1e-10) v = (v / mag) * wrap_pi(2.0 * std::atan2(mag, q.w()));
  return v;
}
inline M3D Rinvleft(const V3D& axis) {
  double theta = axis.norm();
  if (theta < 1e-10) return M3D::Identity();
  double h = theta / 2.0;
  V3D a = axis / theta;
  double s = h * (std::cos(h) / std::sin(h));
  M3D aat;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) aat(i, j) = a(i) * a(j);
  return s * M3D::Identity() + (1.0 - s) * aat - h * skew(a);
}
inline Q4D rpy2Quat(const V3D& rpy) {
  double hy = rpy(2) * 0.5, hp = rpy(1) * 0.5, hr = rpy(0) * 0.5;
  double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
  return Q4D(cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy);
}
inline V3D R2rpy(const M3D& R) {
  V3D rpy;
  rpy(1) = std::atan2(-R(2, 0), std::sqrt(R(2, 1) * R(2, 1) + R(2, 2) * R(2, 2)));
  rpy(0) = std::atan2(R(2, 1) / std::cos(rpy(1)), R(2, 2) / std::cos(rpy(1)));
  rpy(2) = std::atan2(R(1, 0) / std::cos(rpy(1)), R(0, 0) / std::cos(rpy(1)));
  return rpy;
}
inline V3D Q2rpy(const Q4D& q) { return R2rpy(q.toRotationMatrix()); }
inline double deg2rad(double d) { return d * M_PI / 180.0; }
inline double rad2deg(double r) { return r * 180.0 / M_PI; }
// rotation matrix -> quaternion (Eigen's Quaternion(Matrix3) algorithm: Shepperd's method)
inline Q4D R2Quat(const M3D& m) {
  Q4D q;
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w() = 0.5 * t;
    t = 0.5 / t;
    q.x() = (m(2, 1) - m(1, 2)) * t; q.y() = (m(0, 2) - m(2, 0)) * t; q.z() = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q.c[i] = 0.5 * t;
    t = 0.5 / t;
    q.w() = (m(k, j) - m(j, k)) * t;
    q.c[j] = (m(j, i) + m(i, j)) * t;
    q.c[k] = (m(k, i) + m(i, k)) * t;
  }
  return q;
}

}  // namespace math_utils
}  // namespace lins
#endif
