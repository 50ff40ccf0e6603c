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
// GetState() after every message (ros_node.cc:478,515,592,638) pays one device→host copy per message
// (rekf_get_state: one stream synchronisation; the covariance mirror is page-locked so the copy runs at
// PCIe speed), like the by-value copy it pays today.  Writes through the returned references are NOT
// pushed back to the device (the node never writes through them).  A node that only needs the pose and
// the markers should call GetPose() / GetMarkers() instead: 96 bytes instead of n² doubles per message.
//
// Error convention: the reference logs and calls exit(-1) (reflector_ekf_slam.cc:376-377); the C ABI
// returns status codes; this adapter prints rekf_last_error() to stderr and calls std::exit(-1) for
// fatal codes, and keeps the reference's silent behaviour for stale odometry / missing map files.
// The engine's sticky device flags travel with every state read: a landmark-capacity overflow (the
// reference grows without bound, :316-363; the engine's capacity is the constructor's max_landmarks), an
// innovation matrix that lost positive definiteness or a tensor-kernel timeout are reported on stderr
// once and — being unrecoverable divergences from the reference — end the process like the reference's
// own fatal paths.
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
      : handle_(nullptr), mu_stale_(true), sigma_stale_(true), pinned_(nullptr)
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
  ~ReflectorEKFSLAMB200() override
  {
    if (pinned_)
      rekf_host_unregister(handle_, pinned_);
    rekf_destroy(handle_);
  }

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
    out.time = GetLatestTime();   // :99 `result = state_` and the time is never advanced: the reference returns state_.time
    out.mu.resize(n);
    out.sigma.resize(n, n);
    Check(rekf_predict_state(handle_, 0, time, out.mu.data(), n, out.sigma.data(), n), "rekf_predict_state");
    return out;
  }
  Eigen::VectorXd &GetStateVector() override { Refresh(false); return mirror_.mu; }
  Eigen::MatrixXd &GetCoviarance() override { Refresh(true); return mirror_.sigma; }
  double GetLatestTime() override
  {
    double t = 0.;
    Check(rekf_time(handle_, 0, &t), "rekf_time");
    return t;
  }
  State GetState() override
  {
    Refresh(true);
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
  // Node::ReflectorToRosMarkers (ros_node.cc:736-789) evaluated on the device: x, y, angle, x_len, y_len per landmark
  int GetMarkers(std::vector<double> &markers)
  {
    int count = 0;
    Check(rekf_get_markers(handle_, 0, nullptr, 0, &count), "rekf_get_markers");
    markers.resize(5 * static_cast<size_t>(count > 0 ? count : 1));
    if (count > 0)
      Check(rekf_get_markers(handle_, 0, markers.data(), count, &count), "rekf_get_markers");
    return count;
  }
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
  // one rekf_get_state per stale read; a second one when the state has grown since the mirror was sized
  void Refresh(bool want_sigma)
  {
    if (!mu_stale_ && (!want_sigma || !sigma_stale_))
      return;
    for (int attempt = 0; attempt < 2; ++attempt)
    {
      int n = static_cast<int>(mirror_.mu.rows());
      if (n < 3)
        n = Dim();
      ResizeMirror(n, want_sigma);
      int n_now = 0, flags = 0;
      const int rc = rekf_get_state(handle_, 0, n, &mirror_.time, mirror_.mu.data(), want_sigma ? mirror_.sigma.data() : nullptr, n,
                                    &n_now, &flags);
      if (flags != 0)
      {
        std::fprintf(stderr, "[rekf_b200] device flags 0x%x (1: landmark capacity exceeded, 2: S not positive definite, 4: observation "
                             "capacity exceeded, 8: tensor-kernel timeout, 16: solve/Cholesky hand-shake timeout) - the filter no longer follows the reference\n", flags);
        std::exit(-1);
      }
      if (rc == REKF_ERR_CAPACITY && n_now >= 3 && n_now != n)
      {
        ResizeMirror(n_now, want_sigma);
        continue;
      }
      Check(rc, "rekf_get_state");
      mu_stale_ = false;
      if (want_sigma)
        sigma_stale_ = false;
      return;
    }
    Check(REKF_ERR_CAPACITY, "rekf_get_state (state kept growing)");
  }
  void ResizeMirror(int n, bool want_sigma)
  {
    if (mirror_.mu.rows() != n)
    {
      mirror_.mu.resize(n);
      mu_stale_ = sigma_stale_ = true;
    }
    if (want_sigma && (mirror_.sigma.rows() != n || mirror_.sigma.cols() != n))
    {
      if (pinned_)
        rekf_host_unregister(handle_, pinned_);
      mirror_.sigma.resize(n, n);
      sigma_stale_ = true;
      pinned_ = mirror_.sigma.data();
      if (rekf_host_register(handle_, pinned_, sizeof(double) * static_cast<size_t>(n) * n) != REKF_OK)
        pinned_ = nullptr;   // pageable copies still work, only slower
    }
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
  void *pinned_;   // the covariance mirror's buffer while it is page-locked
  bool use_gps_rows_ = false;
  std::vector<float> scratch_;
};
} // namespace ekf

#endif // REFLECTOR_EKF_SLAM_REFLECTOR_EKF_SLAM_B200_H
