// Drives ekf::ReflectorEKFSLAMB200 the way Node::OdometryCallback / ScanCallback drive the reference class
// (reference src/ros_node.cc:627-660, :421-561): one odometry message, one observation, GetState() after
// each.  Reads a stream dumped by tests/test_cpp_adapter.py, writes the final state for comparison with the
// oracle.  Usage: adapter_replay <stream.bin> <out.bin>
// Default: the CI stub of Eigen + the interface (tests/stubs).  -DREKF_ADAPTER_REAL_HEADERS: the reference's own
// ekf_slam_interface.h / sensor_data.h (from /root/reference/include) over the Eigen stand-in of oracle/shim.
#ifndef REKF_ADAPTER_REAL_HEADERS
#define REKF_ADAPTER_STUB_TYPES
#endif
#include "reflector_ekf_slam/reflector_ekf_slam_b200.h"

#include <cstdint>
#include <cstdio>
#include <memory>
#include <vector>

int main(int argc, char **argv)
{
  if (argc < 3) return 2;
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t hdr[4];   // steps, m_stride, N, model
  if (std::fread(hdr, sizeof(int32_t), 4, f) != 4) return 4;
  const int T = hdr[0], m = hdr[1];
  ekf::EKFOptions opt;
  opt.use_imu = false;
  opt.init_time = 0.;
  opt.init_pose = Eigen::Vector3d(0., 0., 0.);
  opt.odom_model = hdr[3] == 0 ? sensor::OdometryModel::DIFF : sensor::OdometryModel::OMNI;
  opt.linear_velocity_cov = 0.05 * 0.05;     // launch/slam.launch:21-23, squared like ros_node.cc:207-237
  opt.angular_velocity_cov = 0.08 * 0.08;
  opt.observation_cov = 0.05 * 0.05;
  std::unique_ptr<ekf::ReflectorEKFSLAMInterface> slam(new ekf::ReflectorEKFSLAMB200(opt, hdr[2], m));
  std::vector<float> xy(2 * static_cast<size_t>(m));
  for (int k = 0; k < T; ++k)
  {
    double od[4], t_obs;
    int32_t cnt;
    if (std::fread(od, sizeof(double), 4, f) != 4 || std::fread(&t_obs, sizeof(double), 1, f) != 1 ||
        std::fread(&cnt, sizeof(int32_t), 1, f) != 1 || std::fread(xy.data(), sizeof(float), xy.size(), f) != xy.size())
      return 5;
    sensor::OdometryData o;
    o.time = od[0];
    o.linear_velocity = Eigen::Vector3d(od[1], od[2], 0.);
    o.angular_velocity = Eigen::Vector3d(0., 0., od[3]);
    slam->HandleOdometryMessage(o);
    (void)slam->GetState();                         // ros_node.cc:638
    sensor::PointCloud cloud;
    for (int i = 0; i < cnt; ++i) cloud.push_back(Eigen::Vector2f(xy[2 * i], xy[2 * i + 1]));
    slam->HandleObservationMessage(sensor::Observation(t_obs, cloud));
    (void)slam->GetState();                         // ros_node.cc:515
  }
  std::fclose(f);
  ekf::State st = slam->GetState();
  FILE *g = std::fopen(argv[2], "wb");
  if (!g) return 6;
  const int32_t n = st.mu.rows();
  std::fwrite(&n, sizeof(int32_t), 1, g);
  std::fwrite(&st.time, sizeof(double), 1, g);
  std::fwrite(st.mu.data(), sizeof(double), static_cast<size_t>(n), g);
  std::fwrite(st.sigma.data(), sizeof(double), static_cast<size_t>(n) * n, g);
  std::fclose(g);
  std::printf("adapter_replay: %d steps, n = %d, pose = %.6f %.6f %.6f\n", T, n, st.mu(0), st.mu(1), st.mu(2));
  return 0;
}
