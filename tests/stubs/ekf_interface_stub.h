// TEST STUB — enough of Eigen and of the reference's value types / interface to compile and exercise
// include/reflector_ekf_slam/reflector_ekf_slam_b200.h where Eigen, ROS and the reference tree are absent
// (this container, the GPU box).  Written for the tests; a real build includes the reference's own
// ekf_slam_interface.h + <Eigen/Dense> instead (REKF_ADAPTER_STUB_TYPES undefined).
#pragma once
#include <cmath>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace Eigen
{
template <typename T, int N>
struct FixedVec
{
  T v[N];
  FixedVec() { for (int i = 0; i < N; ++i) v[i] = T(0); }
  FixedVec(T a, T b) { static_assert(N == 2, "two-argument ctor is for 2-vectors"); v[0] = a; v[1] = b; }
  FixedVec(T a, T b, T c) { static_assert(N == 3, "three-argument ctor is for 3-vectors"); v[0] = a; v[1] = b; v[2] = c; }
  T x() const { return v[0]; }
  T y() const { return v[1]; }
  T z() const { return v[2]; }
  T operator()(int i) const { return v[i]; }
  T &operator()(int i) { return v[i]; }
};
typedef FixedVec<float, 2> Vector2f;
typedef FixedVec<double, 3> Vector3d;

struct Matrix2d
{
  double m[4];   // column-major
  struct Filler
  {
    Matrix2d *self; int k;
    Filler operator,(double v) { self->at(k) = v; return Filler{self, k + 1}; }
  };
  double &at(int k) { return m[(k % 2) * 2 + k / 2]; }   // the comma initialiser fills row by row
  Filler operator<<(double v) { at(0) = v; return Filler{this, 1}; }
  double operator()(int i, int j) const { return m[j * 2 + i]; }
};

class VectorXd
{
public:
  void resize(int n) { d_.assign(static_cast<size_t>(n), 0.0); }
  int rows() const { return static_cast<int>(d_.size()); }
  double *data() { return d_.data(); }
  const double *data() const { return d_.data(); }
  double operator()(int i) const { return d_[i]; }
  double &operator()(int i) { return d_[i]; }
private:
  std::vector<double> d_;
};

class MatrixXd
{
public:
  void resize(int r, int c) { r_ = r; c_ = c; d_.assign(static_cast<size_t>(r) * c, 0.0); }
  int rows() const { return r_; }
  int cols() const { return c_; }
  double *data() { return d_.data(); }
  double operator()(int i, int j) const { return d_[static_cast<size_t>(j) * r_ + i]; }   // column-major
private:
  int r_ = 0, c_ = 0;
  std::vector<double> d_;
};
struct Quaterniond { double w = 1, x = 0, y = 0, z = 0; };
} // namespace Eigen

namespace transform
{
struct Rotation2Dd { double a; double angle() const { return a; } };
class Rigid2d
{
public:
  Rigid2d(double x, double y, double yaw) : t_(x, y, 0.), r_{yaw} {}
  const Eigen::Vector3d &translation() const { return t_; }
  Rotation2Dd rotation() const { return r_; }
private:
  Eigen::Vector3d t_;
  Rotation2Dd r_;
};
} // namespace transform

namespace sensor
{
typedef std::vector<Eigen::Vector2f> PointCloud;
typedef std::vector<Eigen::Matrix2d> PointCloudCoviarance;
class Observation
{
public:
  Observation() : time_(0.) {}
  Observation(const double &time, const PointCloud &cloud) : time_(time), cloud_(cloud) {}
  double time_;
  PointCloud cloud_;
  std::unique_ptr<transform::Rigid2d> gps_pose_;
};
class Map
{
public:
  PointCloud reflector_map_;
  PointCloudCoviarance reflector_map_coviarance_;
};
struct OdometryData
{
  double time;
  Eigen::Vector3d position;
  Eigen::Quaterniond orientation;
  Eigen::Vector3d linear_velocity;
  Eigen::Vector3d angular_velocity;
};
struct ImuData { double time; };
enum OdometryModel { DIFF, OMNI };
} // namespace sensor

namespace ekf
{
struct EKFOptions
{
  bool use_imu;
  double init_time;
  Eigen::Vector3d init_pose;
  std::string map_path;
  sensor::OdometryModel odom_model;
  double linear_velocity_cov, angular_velocity_cov, observation_cov;
};
struct State
{
  double time;
  Eigen::VectorXd mu;
  Eigen::MatrixXd sigma;
};
class ReflectorEKFSLAMInterface
{
public:
  virtual ~ReflectorEKFSLAMInterface() {}
  virtual void HandleOdometryMessage(const sensor::OdometryData &) = 0;
  virtual void HandleImuMessage(const sensor::ImuData &) = 0;
  virtual void HandleObservationMessage(const sensor::Observation &) = 0;
  virtual State PredictState(const double &time) = 0;
  virtual Eigen::VectorXd &GetStateVector() = 0;
  virtual Eigen::MatrixXd &GetCoviarance() = 0;
  virtual double GetLatestTime() = 0;
  virtual State GetState() = 0;
  virtual sensor::Map GetGlobalMap() = 0;
};
} // namespace ekf
