// Host-side mirror of filter::GlobalState / filter::StatePredictor
// (reference: lins/include/KalmanFilter.hpp:35-116 GlobalState, :118-380 StatePredictor).
// 400 Hz scalar IMU propagation stays on the CPU (BASELINE.json north_star: "IMU preintegration ... stay on
// CPU"); the GPU path only consumes state_ / covariance_ through the C-ABI (include/lins_gpu.h) and hands the
// updated pair back through StatePredictor::update.  The reference's global config externs
// (parameters.h:104-153) become the FilterParams struct.  PRODUCT code: never includes oracle/.
#ifndef LINS_HOST_KALMAN_FILTER_HPP_
#define LINS_HOST_KALMAN_FILTER_HPP_

#include <cstring>
#include <vector>

#include "math_utils.hpp"

namespace lins {
namespace filter {

static const double G0 = 9.81;  // parameters.h:62

// exp_port.yaml:29-62 values as defaults
struct FilterParams {
  double acc_n = 70000, gyr_n = 0.1, acc_w = 500, gyr_w = 0.05;
  V3D init_pos_std{0, 0, 0}, init_vel_std{0, 0, 0}, init_att_std{0, 0, 0};
  V3D init_acc_std{0.01, 0.01, 0.02}, init_gyr_std{0.002, 0.002, 0.002};
  V3D init_ba{-0.015774, 0.143237, -0.0263845}, init_bw{-0.00275058, -0.000165954, 0.00262913};
};

class GlobalState {
 public:
  static constexpr unsigned int DIM_OF_STATE_ = 18;
  static constexpr unsigned int DIM_OF_NOISE_ = 12;
  static constexpr unsigned int pos_ = 0, vel_ = 3, att_ = 6, acc_ = 9, gyr_ = 12, gra_ = 15;

  GlobalState() { setIdentity(); }
  GlobalState(const V3D& rn, const V3D& vn, const Q4D& qbn, const V3D& ba, const V3D& bw) {
    setIdentity();
    rn_ = rn; vn_ = vn; qbn_ = qbn; ba_ = ba; bw_ = bw;
  }
  void setIdentity() {
    rn_.setZero(); vn_.setZero(); qbn_.setIdentity(); ba_.setZero(); bw_.setZero();
    gn_ = V3D(0.0, 0.0, -G0);
  }
  void boxPlus(const std::array<double, 18>& xk, GlobalState& out) const {
    GlobalState r = *this;
    for (int i = 0; i < 3; ++i) {
      r.rn_(i) = rn_(i) + xk[pos_ + i]; r.vn_(i) = vn_(i) + xk[vel_ + i];
      r.ba_(i) = ba_(i) + xk[acc_ + i]; r.bw_(i) = bw_(i) + xk[gyr_ + i]; r.gn_(i) = gn_(i) + xk[gra_ + i];
    }
    r.qbn_ = (qbn_ * math_utils::axis2Quat(V3D(xk[att_], xk[att_ + 1], xk[att_ + 2]))).normalized();
    out = r;
  }
  void boxMinus(const GlobalState& in, std::array<double, 18>& xk) const {
    V3D da = math_utils::Quat2axis(in.qbn_.inverse() * qbn_);
    for (int i = 0; i < 3; ++i) {
      xk[pos_ + i] = rn_(i) - in.rn_(i); xk[vel_ + i] = vn_(i) - in.vn_(i); xk[att_ + i] = da(i);
      xk[acc_ + i] = ba_(i) - in.ba_(i); xk[gyr_ + i] = bw_(i) - in.bw_(i); xk[gra_ + i] = gn_(i) - in.gn_(i);
    }
  }
  // C-ABI layout (include/lins_gpu.h): rn vn q(x,y,z,w) ba bw gn
  void toArray(double* s) const {
    for (int i = 0; i < 3; ++i) { s[i] = rn_(i); s[3 + i] = vn_(i); s[10 + i] = ba_(i); s[13 + i] = bw_(i); s[16 + i] = gn_(i); }
    for (int i = 0; i < 4; ++i) s[6 + i] = qbn_.c[i];
  }
  static GlobalState fromArray(const double* s) {
    GlobalState g;
    for (int i = 0; i < 3; ++i) { g.rn_(i) = s[i]; g.vn_(i) = s[3 + i]; g.ba_(i) = s[10 + i]; g.bw_(i) = s[13 + i]; g.gn_(i) = s[16 + i]; }
    for (int i = 0; i < 4; ++i) g.qbn_.c[i] = s[6 + i];
    return g;
  }

  V3D rn_, vn_;
  Q4D qbn_;
  V3D ba_, bw_, gn_;
};

// dense 18x18, column-major like Eigen so covariance_.data() is what the C-ABI takes
struct Cov18 {
  double a[324];
  Cov18() { setZero(); }
  void setZero() { std::memset(a, 0, sizeof(a)); }
  double& operator()(int r, int c) { return a[c * 18 + r]; }
  double operator()(int r, int c) const { return a[c * 18 + r]; }
  const double* data() const { return a; }
  double* data() { return a; }
  void setBlock3(int r0, int c0, const M3D& m) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) (*this)(r0 + i, c0 + j) = m(i, j); }
  M3D block3(int r0, int c0) const { M3D m; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(i, j) = (*this)(r0 + i, c0 + j); return m; }
  void setDiag3(int r0, const V3D& d) { for (int i = 0; i < 3; ++i) (*this)(r0 + i, r0 + i) = d(i); }
};

class StatePredictor {
 public:
  explicit StatePredictor(const FilterParams& p = FilterParams()) : prm_(p) {
    time_ = 0; flag_init_state_ = false; flag_init_imu_ = false;
    reset();
  }

  // KalmanFilter.hpp:125-186
  bool predict(double dt, const V3D& acc, const V3D& gyr, bool update_jacobian_ = true) {
    if (!isInitialized()) return false;
    if (!flag_init_imu_) { flag_init_imu_ = true; acc_last = acc; gyr_last = gyr; }
    GlobalState st = state_;
    V3D un_acc_0 = st.qbn_ * (acc_last - st.ba_) + st.gn_;
    V3D un_gyr = 0.5 * (gyr_last + gyr) - st.bw_;
    Q4D dq = math_utils::axis2Quat(un_gyr * dt);
    st.qbn_ = (st.qbn_ * dq).normalized();
    V3D un_acc_1 = st.qbn_ * (acc - st.ba_) + st.gn_;
    V3D un_acc = 0.5 * (un_acc_0 + un_acc_1);
    st.rn_ = st.rn_ + dt * st.vn_ + 0.5 * dt * dt * un_acc;
    st.vn_ = st.vn_ + dt * un_acc;

    if (update_jacobian_) {
      // Ft (18x18), only these blocks are non-zero (KalmanFilter.hpp:148-160)
      std::vector<double> Ft(324, 0.0), F(324, 0.0), FF(324, 0.0);
      auto at = [](std::vector<double>& m, int r, int c) -> double& { return m[r * 18 + c]; };
      M3D R = st.qbn_.toRotationMatrix();
      M3D va = -(R * math_utils::skew(acc - st.ba_));
      M3D aa = -math_utils::skew(gyr - st.bw_);
      for (int i = 0; i < 3; ++i) {
        at(Ft, GlobalState::pos_ + i, GlobalState::vel_ + i) = 1.0;
        at(Ft, GlobalState::vel_ + i, GlobalState::gra_ + i) = 1.0;
        at(Ft, GlobalState::att_ + i, GlobalState::gyr_ + i) = -1.0;
        for (int j = 0; j < 3; ++j) {
          at(Ft, GlobalState::vel_ + i, GlobalState::att_ + j) = va(i, j);
          at(Ft, GlobalState::vel_ + i, GlobalState::acc_ + j) = -R(i, j);
          at(Ft, GlobalState::att_ + i, GlobalState::att_ + j) = aa(i, j);
        }
      }
      // F_ = I + Ft*dt + 0.5*Ft*Ft*dt*dt
      for (int i = 0; i < 18; ++i)
        for (int k = 0; k < 18; ++k) {
          double f = at(Ft, i, k);
          if (f == 0.0) continue;
          for (int j = 0; j < 18; ++j) at(FF, i, j) += f * at(Ft, k, j);
        }
      for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) at(F, i, j) = (i == j ? 1.0 : 0.0) + at(Ft, i, j) * dt + 0.5 * at(FF, i, j) * dt * dt;
      // Gt*noise*Gt^T with Gt = [vel<-(-R) n_a ; att<-(-I) n_g ; acc<-I n_ba ; gyr<-I n_bg] * dt
      std::vector<double> Q(324, 0.0);
      double dt2 = dt * dt;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += R(i, k) * R(j, k);
          at(Q, GlobalState::vel_ + i, GlobalState::vel_ + j) = s * noise_[0] * dt2;
        }
      for (int i = 0; i < 3; ++i) {
        at(Q, GlobalState::att_ + i, GlobalState::att_ + i) = noise_[1] * dt2;
        at(Q, GlobalState::acc_ + i, GlobalState::acc_ + i) = noise_[2] * dt2;
        at(Q, GlobalState::gyr_ + i, GlobalState::gyr_ + i) = noise_[3] * dt2;
      }
      // covariance_ = F P F^T + Q, then symmetrise
      std::vector<double> FP(324, 0.0), P2(324, 0.0);
      for (int i = 0; i < 18; ++i)
        for (int k = 0; k < 18; ++k) {
          double f = at(F, i, k);
          if (f == 0.0) continue;
          for (int j = 0; j < 18; ++j) FP[i * 18 + j] += f * covariance_(k, j);
        }
      for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) {
          double s = 0;
          for (int k = 0; k < 18; ++k) s += FP[i * 18 + k] * at(F, j, k);
          P2[i * 18 + j] = s + at(Q, i, j);
        }
      for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) covariance_(i, j) = 0.5 * (P2[i * 18 + j] + P2[j * 18 + i]);
    }
    state_ = st;
    time_ += dt;
    acc_last = acc; gyr_last = gyr;
    return true;
  }

  static void calculateRPfromIMU(const V3D& acc, double& roll, double& pitch) {
    pitch = -math_utils::sign(acc.z()) * std::asin(acc.x() / G0);
    roll = math_utils::sign(acc.z()) * std::asin(acc.y() / G0);
  }
  void set(const GlobalState& state) { state_ = state; }
  // KalmanFilter.hpp:195-200
  void update(const GlobalState& state, const Cov18& covariance) { state_ = state; covariance_ = covariance; }

  // KalmanFilter.hpp:210-222
  void initialization(double time, const V3D& rn, const V3D& vn, const Q4D& qbn, const V3D& ba, const V3D& bw,
                      const V3D& acc, const V3D& gyr) {
    state_ = GlobalState(rn, vn, qbn, ba, bw);
    time_ = time; acc_last = acc; gyr_last = gyr;
    flag_init_imu_ = true; flag_init_state_ = true;
    initializeCovariance();
  }
  // KalmanFilter.hpp:234-245 (roll/pitch/yaw default 0)
  void initialization(double time, const V3D& rn, const V3D& vn, const V3D& ba, const V3D& bw, const V3D& acc,
                      const V3D& gyr, double roll = 0.0, double pitch = 0.0, double yaw = 0.0) {
    initialization(time, rn, vn, math_utils::rpy2Quat(V3D(roll, pitch, yaw)), ba, bw, acc, gyr);
  }

  // KalmanFilter.hpp:247-312 (type 0)
  void initializeCovariance() {
    auto sq = [](const V3D& v) { return V3D(v(0) * v(0), v(1) * v(1), v(2) * v(2)); };
    V3D covAtt(std::pow(math_utils::deg2rad(prm_.init_att_std(0)), 2), std::pow(math_utils::deg2rad(prm_.init_att_std(1)), 2),
               std::pow(math_utils::deg2rad(prm_.init_att_std(2)), 2));
    covariance_.setZero();
    covariance_.setDiag3(GlobalState::pos_, sq(prm_.init_pos_std));
    covariance_.setDiag3(GlobalState::vel_, sq(prm_.init_vel_std));
    covariance_.setDiag3(GlobalState::att_, covAtt);
    covariance_.setDiag3(GlobalState::acc_, sq(prm_.init_acc_std));
    covariance_.setDiag3(GlobalState::gyr_, sq(prm_.init_gyr_std));
    covariance_.setDiag3(GlobalState::gra_, V3D(0.01, 0.01, 0.01));
    setNoise();
  }
  void setNoise() {
    const double deg = M_PI / 180.0, dph = deg / 3600.0, dpsh = deg / std::sqrt(3600.0);
    const double ug = (G0 / 1000.0) / 1000.0, ugpsHz = ug / std::sqrt(1.0);
    noise_[0] = std::pow(prm_.acc_n * ug, 2);
    noise_[1] = std::pow(prm_.gyr_n * dph, 2);
    noise_[2] = std::pow(prm_.acc_w * ugpsHz, 2);
    noise_[3] = std::pow(prm_.gyr_w * dpsh, 2);
  }

  // KalmanFilter.hpp:314-354
  void reset(int type = 0) {
    if (type == 0) {
      state_.rn_.setZero();
      state_.vn_ = state_.qbn_.inverse() * state_.vn_;
      state_.qbn_.setIdentity();
      initializeCovariance();
    } else if (type == 1) {
      auto sq = [](const V3D& v) { return V3D(v(0) * v(0), v(1) * v(1), v(2) * v(2)); };
      V3D covAtt(std::pow(math_utils::deg2rad(prm_.init_att_std(0)), 2), std::pow(math_utils::deg2rad(prm_.init_att_std(1)), 2),
                 std::pow(math_utils::deg2rad(prm_.init_att_std(2)), 2));
      M3D vel_cov = covariance_.block3(GlobalState::vel_, GlobalState::vel_);
      M3D acc_cov = covariance_.block3(GlobalState::acc_, GlobalState::acc_);
      M3D gyr_cov = covariance_.block3(GlobalState::gyr_, GlobalState::gyr_);
      M3D gra_cov = covariance_.block3(GlobalState::gra_, GlobalState::gra_);
      M3D Rinv = state_.qbn_.inverse().toRotationMatrix(), R = state_.qbn_.toRotationMatrix();
      covariance_.setZero();
      covariance_.setDiag3(GlobalState::pos_, sq(prm_.init_pos_std));
      covariance_.setBlock3(GlobalState::vel_, GlobalState::vel_, Rinv * vel_cov * R);
      covariance_.setDiag3(GlobalState::att_, covAtt);
      covariance_.setBlock3(GlobalState::acc_, GlobalState::acc_, acc_cov);
      covariance_.setBlock3(GlobalState::gyr_, GlobalState::gyr_, gyr_cov);
      covariance_.setBlock3(GlobalState::gra_, GlobalState::gra_, Rinv * gra_cov * R);
      state_.rn_.setZero();
      state_.vn_ = state_.qbn_.inverse() * state_.vn_;
      state_.qbn_.setIdentity();
      // NB reference order (KalmanFilter.hpp:343-349): qbn_ is already identity when gn_ is "rotated", so
      // gn_ is only re-normalised to 9.81.  Kept as is.
      state_.gn_ = state_.qbn_.inverse() * state_.gn_;
      state_.gn_ = state_.gn_ * 9.81 / state_.gn_.norm();
    }
  }
  inline bool isInitialized() { return flag_init_state_; }

  FilterParams prm_;
  GlobalState state_;
  double time_;
  Cov18 covariance_;
  double noise_[4];  // diagonal of the 12x12 noise_ matrix: (acc_n, gyr_n, acc_w, gyr_w), each x3
  V3D acc_last, gyr_last;
  bool flag_init_state_, flag_init_imu_;
};

}  // namespace filter
}  // namespace lins
#endif
