// reflector_ekf_slam_b200.h — header-only adapter that puts the B200 engine (librekf_b200.so, C ABI in
// include/rekf.h) behind the reference's own interface, ekf::ReflectorEKFSLAMInterface
// (reference include/reflector_ekf_slam/ekf_slam_interface.h:50-67).  C++11, no CUDA or torch types.
//
// Drop-in: replace `ekf::ReflectorEKFSLAM` by `ekf::ReflectorEKFSLAMB200` at the two places the node
// constructs it (reference src/ros_node.cc:436 and :577) and link librekf_b200.so — see INTEGRATION.md
// and patches/ros_node_b200.patch.
//
// The reference class keeps μ and Σ in host Eigen storage and hands out mutable references
// (reflector_ekf_slam.h:25-32).  Here the state lives in HBM; the adapter keeps a host mirror that is
// refreshed lazily — only when a getter is called after the state changed — so a node that reads
// GetState() after every message (ros_node.cc:478,515,592,638) pays one device→host copy per message,
// like the by-value copy it pays today.  Writes through the returned references are NOT pushed back to
// the device (the node never writes through them).
//
// Error convention: the reference logs and calls exit(-1) (reflector_ekf_slam.cc:376-377); the C ABI
// returns status codes; this adapter prints rekf_last_error() to stderr and calls std::exit(-1) for
// fatal codes, and keeps the reference's silent behaviour for stale odometry / missing map files.
#ifndef REFLECTOR_EKF_SLAM_REFLECTOR_EKF_SLAM_B200_H
#define REFLECTOR_EKF_SLAM_REFLECTOR_EKF_SLAM_B200_H

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#ifdef REKF_ADAPTER_STUB_TYPES
#include "ekf_interface_stub.h"   // tests/stubs: minimal Eigen + interface declarations (no Eigen/ROS in CI)
#else
#include "reflector_ekf_slam/ekf_slam_interface.h"
#endif
#include "rekf.h"

namespace ekf
{
class ReflectorEKFSLAMB200 : public ReflectorEKFSLAMInterface
{
public:
  // Same first argument as ReflectorEKFSLAM(const EKFOptions&) (reflector_ekf_slam.cc:6); the capacities
  // are engine-only (the reference grows its Eigen matrices on demand, :320).
  explicit ReflectorEKFSLAMB200(const EKFOptions &options, int max_landmarks = 1024, int max_observations = 128,
                                int device = 0)
      : handle_(nullptr), mu_stale_(true), sigma_stale_(true)
  {
    rekf_options o;
    rekf_default_options(&o);
    o.use_imu = options.use_imu ? 1 : 0;
    o.init_time = options.init_time;
    for (int i = 0; i < 3; ++i)
      o.init_pose[i] = options.init_pose(i);
    o.map_path = options.map_path.c_str();
    o.odom_model = options.odom_model == sensor::OdometryModel::DIFF ? REKF_ODOM_DIFF : REKF_ODOM_OMNI;
    o.linear_velocity_cov = options.linear_velocity_cov;
    o.angular_velocity_cov = options.angular_velocity_cov;
    o.observation_cov = options.observation_cov;
    o.max_landmarks = max_landmarks;
    o.max_observations = max_observations;
    o.device = device;
    Check(rekf_create(&o, &handle_), "rekf_create");
    mirror_.time = options.init_time;
  }
  ReflectorEKFSLAMB200() = delete;
  ReflectorEKFSLAMB200(const ReflectorEKFSLAMB200 &) = delete;
  ReflectorEKFSLAMB200 &operator=(const ReflectorEKFSLAMB200 &) = delete;
  ~ReflectorEKFSLAMB200() override { rekf_destroy(handle_); }

  // reflector_ekf_slam.cc:208-223
  void HandleOdometryMessage(const sensor::OdometryData &odometry) override
  {
    Check(rekf_handle_odometry(handle_, odometry.time, odometry.linear_velocity.x(), odometry.linear_velocity.y(),
                               odometry.angular_velocity.z()),
          "rekf_handle_odometry");
    mu_stale_ = sigma_stale_ = true;
  }
  // reflector_ekf_slam.cc:224-227 (empty in the reference too)
  void HandleImuMessage(const sensor::ImuData &) override {}
  // reflector_ekf_slam.cc:229-368
  void HandleObservationMessage(const sensor::Observation &observation) override
  {
    const int m = static_cast<int>(observation.cloud_.size());
    scratch_.resize(2 * static_cast<size_t>(m > 0 ? m : 1));
    for (int i = 0; i < m; ++i)
    {
      scratch_[2 * i] = observation.cloud_[i].x();
      scratch_[2 * i + 1] = observation.cloud_[i].y();
    }
    double gps[3];
    const double *gps_ptr = nullptr;
    if (observation.gps_pose_)   // reflector_ekf_slam_gps.cc:305: only the GPS variant looks at it
    {
      gps[0] = observation.gps_pose_->translation().x();
      gps[1] = observation.gps_pose_->translation().y();
      gps[2] = observation.gps_pose_->rotation().angle();
      gps_ptr = use_gps_rows_ ? gps : nullptr;
    }
    Check(rekf_handle_observation(handle_, observation.time_, scratch_.data(), m, gps_ptr), "rekf_handle_observation");
    mu_stale_ = sigma_stale_ = true;
  }
  // reflector_ekf_slam.cc:97-152
  State PredictState(const double &time) override
  {
    State out;
    const int n = Dim();
    out.time = time;
    out.mu.resize(n);
    out.sigma.resize(n, n);
    Check(rekf_predict_state(handle_, 0, time, out.mu.data(), n, out.sigma.data(), n), "rekf_predict_state");
    return out;
  }
  Eigen::VectorXd &GetStateVector() override { RefreshMu(); return mirror_.mu; }
  Eigen::MatrixXd &GetCoviarance() override { RefreshSigma(); return mirror_.sigma; }
  double GetLatestTime() override
  {
    double t = 0.;
    Check(rekf_time(handle_, 0, &t), "rekf_time");
    return t;
  }
  State GetState() override
  {
    RefreshMu();
    RefreshSigma();
    mirror_.time = GetLatestTime();
    return mirror_;
  }
  sensor::Map GetGlobalMap() override
  {
    int count = 0;
    Check(rekf_get_map(handle_, nullptr, nullptr, 0, &count), "rekf_get_map");
    std::vector<float> xy(2 * static_cast<size_t>(count > 0 ? count : 1));
    std::vector<double> cov(4 * static_cast<size_t>(count > 0 ? count : 1));
    if (count > 0)
      Check(rekf_get_map(handle_, xy.data(), cov.data(), count, &count), "rekf_get_map");
    sensor::Map map;
    for (int i = 0; i < count; ++i)
    {
      map.reflector_map_.push_back(Eigen::Vector2f(xy[2 * i], xy[2 * i + 1]));
      Eigen::Matrix2d p;
      p << cov[4 * i], cov[4 * i + 1], cov[4 * i + 2], cov[4 * i + 3];
      map.reflector_map_coviarance_.push_back(p);
    }
    return map;
  }

  // ---- engine extras (not part of the reference interface) -----------------------------------------
  // pose + 3x3 block without pulling the whole covariance (what ros_node.cc:802-817 publishes)
  void GetPose(double pose[3], double cov33[9]) { Check(rekf_get_pose(handle_, 0, pose, cov33), "rekf_get_pose"); }
  // Node::SaveReflectorResult (ros_node.cc:75-140) without the host-side State copy
  bool SaveMapTxt(const std::string &filebase) { return rekf_save_map_txt(handle_, 0, filebase.c_str()) == REKF_OK; }
  // feed observation.gps_pose_ as the three pose rows of reflector_ekf_slam_gps.cc:305-340
  void EnableGpsRows(bool on) { use_gps_rows_ = on; }
  rekf_handle *handle() { return handle_; }

private:
  int Dim()
  {
    const int n = rekf_dim(handle_, 0);
    if (n < 0)
      Check(n, "rekf_dim");
    return n;
  }
  void RefreshMu()
  {
    if (!mu_stale_)
      return;
    const int n = Dim();
    mirror_.mu.resize(n);
    int got = 0;
    Check(rekf_get_mu(handle_, 0, mirror_.mu.data(), n, &got), "rekf_get_mu");
    mu_stale_ = false;
  }
  void RefreshSigma()
  {
    if (!sigma_stale_)
      return;
    const int n = Dim();
    mirror_.sigma.resize(n, n);
    Check(rekf_get_sigma(handle_, 0, mirror_.sigma.data(), n), "rekf_get_sigma");   // column-major like MatrixXd
    sigma_stale_ = false;
  }
  void Check(int rc, const char *what)
  {
    if (rc >= 0)
      return;
    std::fprintf(stderr, "[rekf_b200] %s failed (%d): %s\n", what, rc, rekf_last_error(handle_));
    std::exit(-1);   // the reference's convention for unrecoverable input (reflector_ekf_slam.cc:376-377)
  }

  rekf_handle *handle_;
  State mirror_;
  bool mu_stale_, sigma_stale_;
  bool use_gps_rows_ = false;
  std::vector<float> scratch_;
};
} // namespace ekf

#endif // REFLECTOR_EKF_SLAM_REFLECTOR_EKF_SLAM_B200_H
