// sensor_msgs/LaserScan.h — stand-in for the ROS message header (test infrastructure, oracle/): the
// fields of sensor_msgs/LaserScan that laser_reflector_detect.cc reads, with the same names and types.
#ifndef REKF_ORACLE_SENSOR_MSGS_LASERSCAN_SHIM_H
#define REKF_ORACLE_SENSOR_MSGS_LASERSCAN_SHIM_H
#include <memory>
#include <string>
#include <vector>
#include <cstdint>

namespace ros
{
struct Time
{
  uint32_t sec, nsec;
  Time() : sec(0), nsec(0) {}
  Time(uint32_t s, uint32_t ns) : sec(s), nsec(ns) {}
  double toSec() const { return static_cast<double>(sec) + 1e-9 * static_cast<double>(nsec); }
};
} // namespace ros

namespace std_msgs
{
struct Header
{
  uint32_t seq;
  ros::Time stamp;
  std::string frame_id;
  Header() : seq(0) {}
};
} // namespace std_msgs

namespace sensor_msgs
{
struct LaserScan
{
  std_msgs::Header header;
  float angle_min, angle_max, angle_increment, time_increment, scan_time, range_min, range_max;
  std::vector<float> ranges, intensities;
};
typedef std::shared_ptr<LaserScan> LaserScanPtr;
typedef std::shared_ptr<const LaserScan> LaserScanConstPtr;
} // namespace sensor_msgs
#endif
