// tools/synth/lins_sequence.cpp — offline sequence driver: what LinsFusion does between the
// ROS callbacks and the estimator (lins/src/lib/Estimator.cpp:204-252): for every lidar scan, feed the buffered
// IMU samples through processImu, then the segmented cloud through processPCL.  Inputs are synthetic (a smooth
// trajectory through the seeded world of lins_synth.cpp; IMU at 400 Hz; raw scans pushed through the restated
// image projection).  The estimator is the C++ shim fusion::StateEstimator whose hot seams run on the GPU through
// the C-ABI, so this file links against liblins_gpu.so.  Every performIESKF call's inputs / outputs are recorded
// so that tests can replay them through the CPU oracle.  Never includes anything from oracle/.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lins_synth.cpp"

#include "../../lins---lidar-inertial-slam_b200/csrc/host/state_estimator.hpp"

using lins::fusion::StateEstimator;

namespace {

struct SeqRecord {
  // one recorded performIESKF unit per RUNNING scan
  std::vector<lins_point> surfFlat, cornerSharp, surfLessFlat, cornerLessSharp;
  std::vector<int32_t> offSF{0}, offCS{0}, offSL{0}, offCL{0};
  std::vector<double> state_in, cov_in, state_out, cov_out;
  std::vector<int32_t> iters, flags, scan_index;
  std::vector<double> rel_true;  // true relative pose of that scan: t (3) + q (x,y,z,w)
  std::vector<int32_t> status;   // estimator status after each scan
  std::vector<double> global_est, global_true;  // per scan: position (3) + quaternion (x,y,z,w)
};

void append_cloud(std::vector<lins_point>& dst, std::vector<int32_t>& off, const Cloud& c) {
  dst.insert(dst.end(), c.points.begin(), c.points.end());
  off.push_back((int32_t)dst.size());
}

struct Traj {  // smooth body-frame velocity / yaw-rate profile
  double v0, va, vw, w0, wa, ww;
  V3D vbody(double t) const { return V3D(v0 + va * std::sin(vw * t), 0.1 * std::sin(0.7 * vw * t), 0.0); }
  V3D vdot(double t) const { return V3D(va * vw * std::cos(vw * t), 0.07 * vw * std::cos(0.7 * vw * t), 0.0); }
  V3D wbody(double t) const { return V3D(0.0, 0.0, w0 + wa * std::sin(ww * t)); }
};

}  // namespace

extern "C" {

void* lins_seq_run(const lins_synth_cfg* cfg, uint64_t seed, int n_scans, int device) {
  SeqRecord* rec = new SeqRecord();
  Rng rng(seed);
  LidarModel lm = cfg->lidar == 1 ? LidarModel::dense64() : LidarModel::vlp16();
  World w = make_world(rng, cfg->world);
  const double T = lm.scan_period;
  const int N = lm.scan_num, nimu = 40;
  const double dt = T / nimu;
  Traj tr{rng.uni(1.0, 0.6 * cfg->v_max), rng.uni(0.2, 1.0), rng.uni(0.5, 1.5), rng.uni(-0.5, 0.5) * cfg->w_max, 0.5 * cfg->w_max, rng.uni(0.3, 1.0)};
  Pose P;
  P.R = math_utils::rpy2Quat(V3D(0.0, 0.0, rng.uni(-M_PI, M_PI))).toRotationMatrix();
  P.p = V3D(rng.uni(-3, 3), rng.uni(-3, 3), 1.5);
  const V3D g_w(0, 0, -filter::G0);
  const V3D ba(0.0, 0.0, 0.0), bw(0.0, 0.0, 0.0);  // the shim initialises its biases from FilterParams

  lins::fusion::EstimatorParams ep;
  ep.lidar = lm;
  ep.filter.init_ba = V3D(0, 0, 0);
  ep.filter.init_bw = V3D(0, 0, 0);
  StateEstimator est(ep, device);
  ImageProjection ip(lm);

  const bool verbose = std::getenv("LINS_SEQ_VERBOSE") != nullptr;
  double t = 0.0;
  for (int k = 0; k < n_scans; ++k) {
    if (verbose) std::fprintf(stderr, "[lins_seq] simulating scan %d\n", k);
    // ---- simulate one sweep: poses at the 40 IMU instants, IMU samples, 1800 firings ---------------------------
    std::vector<Pose> poses(nimu + 1);
    std::vector<V3D> accs(nimu + 1), gyrs(nimu + 1);
    poses[0] = P;
    const Pose Pstart = P;
    for (int i = 0; i <= nimu; ++i) {
      const double ti = t + i * dt;
      const V3D wb = tr.wbody(ti), vb = tr.vbody(ti);
      accs[i] = cross(wb, vb) + tr.vdot(ti) - poses[i].R.transpose() * g_w + ba + V3D(0.02 * rng.gauss(), 0.02 * rng.gauss(), 0.02 * rng.gauss());
      gyrs[i] = wb + bw + V3D(5e-4 * rng.gauss(), 5e-4 * rng.gauss(), 5e-4 * rng.gauss());
      if (i < nimu) {  // midpoint step of the true motion
        const double tm = ti + 0.5 * dt;
        poses[i + 1].p = poses[i].p + dt * (poses[i].R * math_utils::axis2Quat(0.5 * dt * tr.wbody(tm)).toRotationMatrix() * tr.vbody(tm));
        poses[i + 1].R = poses[i].R * math_utils::axis2Quat(dt * tr.wbody(tm)).toRotationMatrix();
      }
    }
    Cloud raw;
    for (int f = 0; f < N; ++f) {
      const double s = (double)f / N * nimu;
      const int i0 = std::min((int)s, nimu - 1);
      const double a = s - i0;
      Pose Pf;
      Pf.p = poses[i0].p + a * (poses[i0 + 1].p - poses[i0].p);
      Pf.R = poses[i0].R * math_utils::axis2Quat(a * dt * tr.wbody(t + (i0 + 0.5 * a) * dt)).toRotationMatrix();
      const double ori = -M_PI + (f + 0.25) * (2.0 * M_PI / N);
      for (int r = 0; r < lm.line_num; ++r) {
        const double el = (-(double)(lm.ang_bottom - 0.1f) + r * (double)lm.ang_res_y) * M_PI / 180.0;
        const V3D ds(std::cos(el) * std::cos(ori), -std::cos(el) * std::sin(ori), std::sin(el));
        double rg = raycast(w, Pf.p, Pf.R * ds);
        if (!(rg < 100.0)) continue;
        rg += cfg->range_noise * rng.gauss();
        raw.push_back(makePoint((float)(rg * ds.x()), (float)(rg * ds.y()), (float)(rg * ds.z()), 0.f));
      }
    }
    P = poses[nimu];
    t += T;
    // ---- what LinsFusion::processPointClouds does (Estimator.cpp:204-252) -----------------------------------------
    ip.process(raw);
    for (int i = 1; i <= nimu; ++i) est.processImu(dt, accs[i], gyrs[i]);
    const bool will_run = est.status_ == StateEstimator::STATUS_RUNNING;
    double s_in[19];
    filter::Cov18 P_in;
    std::vector<lins_point> mapS, mapC;
    if (will_run) {
      est.filter_->state_.toArray(s_in);
      P_in = est.filter_->covariance_;
      mapS = est.scan_last_->surfPointsLessFlat_.points;
      mapC = est.scan_last_->cornerPointsLessSharp_.points;
    }
    est.last_report_.iters = 0;
    est.processPCL(t, lins::sensor_utils::Imu(t, accs[nimu], gyrs[nimu]), ip.segmentedCloud, ip.segMsg, ip.outlierCloud);
    rec->status.push_back((int)est.status_);
    if (verbose) std::fprintf(stderr, "[lins_seq] scan %d: status %d, IESKF iterations %d, %zu segmented points\n", k, (int)est.status_, (int)est.last_report_.iters, ip.segmentedCloud.size());
    if (will_run && est.last_report_.iters > 0) {
      // processScan swapped the scans: scan_last_ now IS the scan whose features were the queries
      append_cloud(rec->surfFlat, rec->offSF, est.scan_last_->surfPointsFlat_);
      append_cloud(rec->cornerSharp, rec->offCS, est.scan_last_->cornerPointsSharp_);
      Cloud ms, mc; ms.points = mapS; mc.points = mapC;
      append_cloud(rec->surfLessFlat, rec->offSL, ms);
      append_cloud(rec->cornerLessSharp, rec->offCL, mc);
      rec->state_in.insert(rec->state_in.end(), s_in, s_in + 19);
      rec->cov_in.insert(rec->cov_in.end(), P_in.data(), P_in.data() + 324);
      // filter_->state_ after performIESKF but BEFORE reset(1) is linState_ (not diverged) — reset(1) zeroes rn/qbn,
      // so take the relative pose from linState_ and the remaining blocks from the filter
      double s_out[19];
      lins::filter::GlobalState so = est.linState_;
      so.toArray(s_out);
      rec->state_out.insert(rec->state_out.end(), s_out, s_out + 19);
      rec->iters.push_back(est.last_report_.iters);
      rec->flags.push_back((est.last_report_.converged ? 1 : 0) | (est.last_report_.diverged ? 2 : 0) | (est.last_report_.has_nan ? 4 : 0));
      rec->scan_index.push_back(k);
      // true relative pose of this sweep
      M3D Rrel = Pstart.R.transpose() * P.R;
      V3D trel = Pstart.R.transpose() * (P.p - Pstart.p);
      Q4D qrel = math_utils::R2Quat(Rrel);
      const double tv[7] = {trel.x(), trel.y(), trel.z(), qrel.x(), qrel.y(), qrel.z(), qrel.w()};
      rec->rel_true.insert(rec->rel_true.end(), tv, tv + 7);
    }
    const double ge[7] = {est.globalState_.rn_.x(), est.globalState_.rn_.y(), est.globalState_.rn_.z(), est.globalState_.qbn_.x(),
                          est.globalState_.qbn_.y(), est.globalState_.qbn_.z(), est.globalState_.qbn_.w()};
    rec->global_est.insert(rec->global_est.end(), ge, ge + 7);
    Q4D qt = math_utils::R2Quat(P.R);
    const double gt[7] = {P.p.x(), P.p.y(), P.p.z(), qt.x(), qt.y(), qt.z(), qt.w()};
    rec->global_true.insert(rec->global_true.end(), gt, gt + 7);
  }
  return rec;
}

void lins_seq_destroy(void* h) { delete static_cast<SeqRecord*>(h); }
int lins_seq_num_units(void* h) { return (int)static_cast<SeqRecord*>(h)->iters.size(); }
int lins_seq_num_scans(void* h) { return (int)static_cast<SeqRecord*>(h)->status.size(); }
void lins_seq_desc(void* h, lins_batch_desc* d) {
  SeqRecord* r = static_cast<SeqRecord*>(h);
  d->n_scans = (int32_t)r->iters.size();
  d->surf_flat = r->surfFlat.data(); d->surf_flat_off = r->offSF.data();
  d->corner_sharp = r->cornerSharp.data(); d->corner_sharp_off = r->offCS.data();
  d->surf_less_flat = r->surfLessFlat.data(); d->surf_less_flat_off = r->offSL.data();
  d->corner_less_sharp = r->cornerLessSharp.data(); d->corner_less_sharp_off = r->offCL.data();
  d->state_in = r->state_in.data(); d->cov_in = r->cov_in.data();
}
const double* lins_seq_array(void* h, int which) {
  SeqRecord* r = static_cast<SeqRecord*>(h);
  switch (which) {
    case 0: return r->state_out.data();
    case 1: return r->rel_true.data();
    case 2: return r->global_est.data();
    case 3: return r->global_true.data();
  }
  return nullptr;
}
const int32_t* lins_seq_ints(void* h, int which) {
  SeqRecord* r = static_cast<SeqRecord*>(h);
  switch (which) {
    case 0: return r->iters.data();
    case 1: return r->flags.data();
    case 2: return r->scan_index.data();
    case 3: return r->status.data();
  }
  return nullptr;
}

}  // extern "C"
