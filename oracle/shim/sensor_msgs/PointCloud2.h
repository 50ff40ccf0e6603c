// sensor_msgs/PointCloud2.h — stand-in (test infrastructure, oracle/): only named by the detector interface.
#ifndef REKF_ORACLE_SENSOR_MSGS_POINTCLOUD2_SHIM_H
#define REKF_ORACLE_SENSOR_MSGS_POINTCLOUD2_SHIM_H
#include <memory>
#include "sensor_msgs/LaserScan.h"
namespace sensor_msgs
{
struct PointCloud2
{
  std_msgs::Header header;
};
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
} // namespace sensor_msgs
#endif
