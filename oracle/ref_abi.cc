// ref_abi.cc — exposes the reference's OWN classes, ekf::ReflectorEKFSLAM and ekf::ReflectorEKFSLAMGPS, compiled
// UNMODIFIED from /root/reference/src/reflector_ekf_slam/reflector_ekf_slam{,_gps}.cc (against oracle/shim's
// stand-ins for Eigen and glog), through the same C ABI as the C restatement (rekf_oracle.h).
// TEST INFRASTRUCTURE: this is what pins the restatement and the CUDA engine to the reference (oracle/_ref/).
// Built by `make -C oracle ref`; nothing under reflector_ekf_slam_b200/ links or loads it.
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include <Eigen/Dense>
#include <glog/logging.h>

// The checker needs the private association routine and the members a warm start overwrites
// (ReflectorMatch, state_, vt_, map_).  The reference's own translation units are compiled without this.
#define private public
#include "reflector_ekf_slam/reflector_ekf_slam.h"
#include "reflector_ekf_slam/reflector_ekf_slam_gps.h"
#undef private

#include "rekf_oracle.h"

namespace
{
enum { ALGEBRA_GPS_CLASS = 0x100 };

struct Ref
{
  std::unique_ptr<ekf::ReflectorEKFSLAM> plain;
  std::unique_ptr<ekf::ReflectorEKFSLAMGPS> gps;
  ekf::ReflectorMatchResult last_match;
  int map_loader;
  bool track_matches = true;

  ekf::ReflectorEKFSLAMInterface *iface() { return plain ? static_cast<ekf::ReflectorEKFSLAMInterface *>(plain.get()) : gps.get(); }
  ekf::State &state() { return plain ? plain->state_ : gps->state_; }
  Eigen::Vector3d &vt() { return plain ? plain->vt_ : gps->vt_; }
  sensor::Map &map() { return plain ? plain->map_ : gps->map_; }
};

sensor::Observation make_observation(double time, const float *xy, int m)
{
  sensor::PointCloud cloud;
  for (int i = 0; i < m; ++i) cloud.push_back(Eigen::Vector2f(xy[2 * i], xy[2 * i + 1]));
  return sensor::Observation(time, cloud);
}

template <class EKF> ekf::ReflectorMatchResult peek_match(const EKF &live, const sensor::Observation &obs)
{
  ekf::EKFOptions o = live.options_;
  o.map_path.clear();
  EKF peek(o);
  peek.vt_ = live.vt_;
  peek.map_ = live.map_;
  peek.state_.time = live.state_.time;
  peek.state_.mu = Eigen::Vector3d(live.state_.mu(0), live.state_.mu(1), live.state_.mu(2));
  peek.Predict(obs.time_ - live.state_.time); // :232
  const Eigen::Vector3d pose(peek.state_.mu(0), peek.state_.mu(1), peek.state_.mu(2));
  const Eigen::Index n = live.state_.mu.rows();
  peek.state_.mu = live.state_.mu;
  peek.state_.mu.topRows(3) = pose;
  peek.state_.sigma = Eigen::MatrixXd::Zero(n, n); // only ever read as an unused 2x2 block (:432)
  return peek.ReflectorMatch(obs);
}
} // namespace

struct rekf_oracle
{
  Ref r;
};

extern "C" {

rekf_oracle *oracle_create(const rekf_options *opts, int algebra)
{
  ekf::EKFOptions o;
  o.use_imu = opts->use_imu != 0;
  o.init_time = opts->init_time;
  o.init_pose = Eigen::Vector3d(opts->init_pose[0], opts->init_pose[1], opts->init_pose[2]);
  o.map_path = opts->map_path ? opts->map_path : "";
  o.odom_model = opts->odom_model == REKF_ODOM_OMNI ? sensor::OdometryModel::OMNI : sensor::OdometryModel::DIFF;
  o.linear_velocity_cov = opts->linear_velocity_cov;
  o.angular_velocity_cov = opts->angular_velocity_cov;
  o.observation_cov = opts->observation_cov;
  rekf_oracle *h = new rekf_oracle;
  h->r.map_loader = opts->map_loader;
  if (algebra & ALGEBRA_GPS_CLASS)
    h->r.gps.reset(new ekf::ReflectorEKFSLAMGPS(o));
  else
    h->r.plain.reset(new ekf::ReflectorEKFSLAM(o));
  return h;
}

void oracle_destroy(rekf_oracle *h) { delete h; }

void oracle_handle_odometry(rekf_oracle *h, double time, double vx, double vy, double wz)
{
  sensor::OdometryData d;
  d.time = time;
  d.position = Eigen::Vector3d::Zero();
  d.linear_velocity = Eigen::Vector3d(vx, vy, 0.);
  d.angular_velocity = Eigen::Vector3d(0., 0., wz);
  h->r.iface()->HandleOdometryMessage(d);
}

void oracle_handle_observation(rekf_oracle *h, double time, const float *xy, int m, const double *gps_pose)
{
  sensor::Observation obs = make_observation(time, xy, m);
  if (gps_pose)
    obs.gps_pose_.reset(new transform::Rigid2d(Eigen::Vector2d(gps_pose[0], gps_pose[1]), gps_pose[2]));
  // The association the update is about to use, for the match-result getter.  ReflectorMatch is a pure
  // function of (mu after Predict, map, cloud) and Predict's mean update touches only the pose, so run the
  // reference's own Predict on a 3-state look-ahead object, give it the landmarks, and call ReflectorMatch.
  h->r.last_match = ekf::ReflectorMatchResult();
  if (m > 0 && h->r.track_matches)
  {
    if (h->r.plain)
      h->r.last_match = peek_match(*h->r.plain, obs);
    else
      h->r.last_match = peek_match(*h->r.gps, obs);
  }
  h->r.iface()->HandleObservationMessage(obs);
}

int oracle_dim(const rekf_oracle *h) { return static_cast<int>(const_cast<rekf_oracle *>(h)->r.state().mu.rows()); }
double oracle_time(const rekf_oracle *h) { return const_cast<rekf_oracle *>(h)->r.iface()->GetLatestTime(); }
const double *oracle_mu(const rekf_oracle *h) { return const_cast<rekf_oracle *>(h)->r.iface()->GetStateVector().data(); }
const double *oracle_sigma(const rekf_oracle *h) { return const_cast<rekf_oracle *>(h)->r.iface()->GetCoviarance().data(); }

void oracle_get_match_result(const rekf_oracle *h, int *state_pairs, int *n_state, int *map_pairs, int *n_map,
                             int *new_ids, int *n_new, int cap)
{
  const ekf::ReflectorMatchResult &m = h->r.last_match;
  const int ns = static_cast<int>(m.state_obs_match_ids.size()), nm = static_cast<int>(m.map_obs_match_ids.size()),
            nn = static_cast<int>(m.new_ids.size());
  if (n_state) *n_state = ns;
  if (n_map) *n_map = nm;
  if (n_new) *n_new = nn;
  for (int i = 0; state_pairs && i < ns && i < cap; ++i)
    state_pairs[2 * i] = m.state_obs_match_ids[i].first, state_pairs[2 * i + 1] = m.state_obs_match_ids[i].second;
  for (int i = 0; map_pairs && i < nm && i < cap; ++i)
    map_pairs[2 * i] = m.map_obs_match_ids[i].first, map_pairs[2 * i + 1] = m.map_obs_match_ids[i].second;
  for (int i = 0; new_ids && i < nn && i < cap; ++i) new_ids[i] = m.new_ids[i];
}

void oracle_predict_state(const rekf_oracle *h, double time, double *mu, double *sigma)
{
  const ekf::State s = const_cast<rekf_oracle *>(h)->r.iface()->PredictState(time);
  std::memcpy(mu, s.mu.data(), sizeof(double) * static_cast<size_t>(s.mu.rows()));
  if (sigma) std::memcpy(sigma, s.sigma.data(), sizeof(double) * static_cast<size_t>(s.sigma.rows() * s.sigma.cols()));
}

void oracle_set_state(rekf_oracle *h, double time, const double vt[3], const double *mu, int n, const double *sigma, int ld)
{
  ekf::State &s = h->r.state();
  s.time = time;
  s.mu.resize(n);
  s.sigma.resize(n, n);
  for (int i = 0; i < n; ++i) s.mu(i) = mu[i];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) s.sigma(i, j) = sigma[static_cast<size_t>(j) * ld + i];
  h->r.vt() = Eigen::Vector3d(vt[0], vt[1], vt[2]);
}

void oracle_set_map(rekf_oracle *h, const float *xy, const double *cov2x2, int count)
{
  sensor::Map &m = h->r.map();
  m.reflector_map_.clear();
  m.reflector_map_coviarance_.clear();
  for (int i = 0; i < count; ++i)
  {
    m.reflector_map_.push_back(Eigen::Vector2f(xy[2 * i], xy[2 * i + 1]));
    Eigen::Matrix2d p;
    p << cov2x2[4 * i], cov2x2[4 * i + 1], cov2x2[4 * i + 2], cov2x2[4 * i + 3];
    m.reflector_map_coviarance_.push_back(p);
  }
}

int oracle_get_map(const rekf_oracle *h, float *xy, double *cov2x2, int cap)
{
  const sensor::Map m = const_cast<rekf_oracle *>(h)->r.iface()->GetGlobalMap();
  const int count = static_cast<int>(m.reflector_map_.size());
  for (int i = 0; i < count && i < cap; ++i)
  {
    if (xy) xy[2 * i] = m.reflector_map_[i].x(), xy[2 * i + 1] = m.reflector_map_[i].y();
    if (cov2x2 && i < static_cast<int>(m.reflector_map_coviarance_.size()))
    {
      const Eigen::Matrix2d &p = m.reflector_map_coviarance_[i];
      cov2x2[4 * i] = p(0, 0), cov2x2[4 * i + 1] = p(0, 1), cov2x2[4 * i + 2] = p(1, 0), cov2x2[4 * i + 3] = p(1, 1);
    }
  }
  return count;
}

// LoadMapFromTxtFile is private and only reachable through the constructor (:36); calling it directly keeps
// the loaded object the reference's own.
void oracle_load_map_txt(rekf_oracle *h, const char *path)
{
  if (h->r.plain)
    h->r.plain->LoadMapFromTxtFile(path ? path : "");
  else
    h->r.gps->LoadMapFromTxtFile(path ? path : "");
}

// Node::SaveReflectorResult lives in ros_node.cc (ROS-bound, not compilable here): not provided by _ref.
int oracle_save_map_txt(const rekf_oracle *, const char *) { return -1; }

int oracle_ref_is_reference(void) { return 1; }
// timing runs switch the look-ahead association (an extra O(n^2) per frame) off
void oracle_ref_track_matches(rekf_oracle *h, int on) { h->r.track_matches = on != 0; }
}
