// detect_abi.cc — C ABI around the reference's OWN reflector_detect::LaserReflectorDetect (+ PoseExtrapolator),
// compiled UNMODIFIED from /root/reference/src/reflector_detect/laser/*.cc against oracle/shim's stand-ins for
// Eigen, glog and sensor_msgs.  TEST INFRASTRUCTURE (oracle/_ref/libdetect_ref.so): pins the replay/ front-end
// (the ROS-free restatement of the detector) to the reference's code on real bag scans.
#include <cmath>
#include <memory>
#include <vector>

#include <Eigen/Dense>
#include <glog/logging.h>

#include "reflector_detect/laser/laser_reflector_detect.h"

struct detect_ref
{
  std::unique_ptr<reflector_detect::LaserReflectorDetect> d;
};

extern "C" {

// options as Node::LoadOptions fills them (ros_node.cc:240-283): intensity_min, reflector_min_length,
// reflector_length_error, range_min, range_max; sensor_to_base_link = (x, y, yaw)
detect_ref *detect_create(double intensity_min, double reflector_min_length, double reflector_length_error,
                          float range_min, float range_max, const double sensor_to_base_link[3])
{
  reflector_detect::ReflectorDetectOptions o;
  o.intensity_min = intensity_min;
  o.reflector_min_length = reflector_min_length;
  o.reflector_length_error = reflector_length_error;
  o.range_min = range_min;
  o.range_max = range_max;
  detect_ref *h = new detect_ref;
  h->d.reset(new reflector_detect::LaserReflectorDetect(o));
  // transform::RollPitchYaw(0, 0, yaw) (ros_node.cc:278) = rotation about z
  const double yaw = sensor_to_base_link[2];
  h->d->SetSensorToBaseLinkTransform(transform::Rigid3d(
      Eigen::Vector3d(sensor_to_base_link[0], sensor_to_base_link[1], 0.),
      Eigen::Quaterniond(std::cos(yaw / 2), 0., 0., std::sin(yaw / 2))));
  return h;
}

void detect_destroy(detect_ref *h) { delete h; }

// Node::ToOdometryData (ros_node.cc:662-680) then LaserReflectorDetect::HandleOdometryData
void detect_handle_odometry(detect_ref *h, double time, const double position[3], const double quat_wxyz[4],
                            const double linear[3], const double angular[3])
{
  sensor::OdometryData d;
  d.time = time;
  d.position = Eigen::Vector3d(position[0], position[1], position[2]);
  d.orientation = Eigen::Quaterniond(quat_wxyz[0], quat_wxyz[1], quat_wxyz[2], quat_wxyz[3]);
  d.linear_velocity = Eigen::Vector3d(linear[0], linear[1], linear[2]);
  d.angular_velocity = Eigen::Vector3d(angular[0], angular[1], angular[2]);
  h->d->HandleOdometryData(d);
}

// LaserReflectorDetect::HandleLaserScan.  stamp as (sec, nsec) like ros::Time.  Returns the number of reflector
// centres (written to xy_out, at most cap pairs) and the observation time.
int detect_handle_scan(detect_ref *h, unsigned sec, unsigned nsec, float angle_min, float angle_max,
                       float angle_increment, float time_increment, float scan_time, float range_min,
                       float range_max, const float *ranges, const float *intensities, int count,
                       double *time_out, float *xy_out, int cap)
{
  std::shared_ptr<sensor_msgs::LaserScan> msg(new sensor_msgs::LaserScan);
  msg->header.stamp = ros::Time(sec, nsec);
  msg->angle_min = angle_min, msg->angle_max = angle_max, msg->angle_increment = angle_increment;
  msg->time_increment = time_increment, msg->scan_time = scan_time;
  msg->range_min = range_min, msg->range_max = range_max;
  msg->ranges.assign(ranges, ranges + count);
  msg->intensities.assign(intensities, intensities + count);
  const sensor::Observation obs = h->d->HandleLaserScan(msg);
  if (time_out) *time_out = obs.time_;
  const int n = static_cast<int>(obs.cloud_.size());
  for (int i = 0; i < n && i < cap; ++i) xy_out[2 * i] = obs.cloud_[i].x(), xy_out[2 * i + 1] = obs.cloud_[i].y();
  return n;
}

// the motion-corrected scan the detector leaves for the grid mapper (GetRangeData): returns count
int detect_range_returns(detect_ref *h, float *xy_out, int cap)
{
  const sensor::RangeData r = h->d->GetRangeData();
  const int n = static_cast<int>(r.returns.size());
  for (int i = 0; i < n && i < cap; ++i) xy_out[2 * i] = r.returns[i].x(), xy_out[2 * i + 1] = r.returns[i].y();
  return n;
}
}
