// adapter_bench.cc — the drop-in path measured: ekf::ReflectorEKFSLAMB200 (the C++11 adapter over the C ABI) driven
// exactly like Node::OdometryCallback / ScanCallback drive the reference class (reference src/ros_node.cc:627-660,
// :421-561): HandleOdometryMessage, GetState(), HandleObservationMessage, GetState() — a full by-value State (mu and the
// n x n covariance) after EVERY message, host buffers in, host buffers out.  Then the same stream with GetPose() in place
// of GetState() (what a node that publishes pose + markers actually needs: 96 bytes per message instead of n² doubles).
// A third pass reads the state through the by-REFERENCE getters (GetStateVector() / GetCoviarance(): the refreshed, page-locked
// mirror, no by-value copy): what the full-covariance read costs without the interface's copy.
// bench.py builds and runs this (`e2e_adapter`).  Usage: adapter_bench <stream.bin> <n_build> <k_state> <k_pose>
// stream.bin: int32 {steps, m_stride, N, model}, then per step: double od[4], double t_obs, int32 cnt, float xy[2*m_stride].
#define REKF_ADAPTER_STUB_TYPES
#include "reflector_ekf_slam/reflector_ekf_slam_b200.h"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

struct Step { double od[4]; double t_obs; int32_t cnt; std::vector<float> xy; };

int main(int argc, char **argv)
{
  if (argc < 5) return 2;
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t hdr[4];
  if (std::fread(hdr, sizeof(int32_t), 4, f) != 4) return 4;
  const int T = hdr[0], m = hdr[1];
  const int n_build = std::atoi(argv[2]), k_state = std::atoi(argv[3]), k_pose = std::atoi(argv[4]);
  if (n_build + 2 * k_state + k_pose + 2 > T) return 5;
  std::vector<Step> steps(static_cast<size_t>(T));
  for (int k = 0; k < T; ++k)
  {
    Step &s = steps[k];
    s.xy.resize(2 * static_cast<size_t>(m));
    if (std::fread(s.od, sizeof(double), 4, f) != 4 || std::fread(&s.t_obs, sizeof(double), 1, f) != 1 ||
        std::fread(&s.cnt, sizeof(int32_t), 1, f) != 1 || std::fread(s.xy.data(), sizeof(float), s.xy.size(), f) != s.xy.size())
      return 6;
  }
  std::fclose(f);
  ekf::EKFOptions opt;
  opt.use_imu = false;
  opt.init_time = 0.;
  opt.init_pose = Eigen::Vector3d(0., 0., 0.);
  opt.odom_model = hdr[3] == 0 ? sensor::OdometryModel::DIFF : sensor::OdometryModel::OMNI;
  opt.linear_velocity_cov = 0.05 * 0.05;     // launch/slam.launch:21-23, squared like ros_node.cc:207-237
  opt.angular_velocity_cov = 0.08 * 0.08;
  opt.observation_cov = 0.05 * 0.05;
  ekf::ReflectorEKFSLAMB200 *b200 = new ekf::ReflectorEKFSLAMB200(opt, hdr[2], m);
  std::unique_ptr<ekf::ReflectorEKFSLAMInterface> slam(b200);
  double sink = 0.;
  auto run = [&](int k, int mode) {   // mode 1: GetState() by value, 0: GetPose(), 2: by-reference getters
    const bool full_state = mode == 1;
    const Step &s = steps[k];
    sensor::OdometryData o;
    o.time = s.od[0];
    o.linear_velocity = Eigen::Vector3d(s.od[1], s.od[2], 0.);
    o.angular_velocity = Eigen::Vector3d(0., 0., s.od[3]);
    double pose[3], cov[9];
    slam->HandleOdometryMessage(o);
    if (full_state) { ekf::State st = slam->GetState(); sink += st.mu(0) + st.sigma(0, 0); }   // ros_node.cc:638
    else if (mode == 2) { sink += slam->GetStateVector()(0) + slam->GetCoviarance()(0, 0); }
    else { b200->GetPose(pose, cov); sink += pose[0] + cov[0]; }
    sensor::PointCloud cloud;
    for (int i = 0; i < s.cnt; ++i) cloud.push_back(Eigen::Vector2f(s.xy[2 * i], s.xy[2 * i + 1]));
    slam->HandleObservationMessage(sensor::Observation(s.t_obs, cloud));
    if (full_state) { ekf::State st = slam->GetState(); sink += st.mu(0) + st.sigma(0, 0); }   // ros_node.cc:515
    else if (mode == 2) { sink += slam->GetStateVector()(0) + slam->GetCoviarance()(0, 0); }
    else { b200->GetPose(pose, cov); sink += pose[0] + cov[0]; }
  };
  int k = 0;
  for (; k < n_build; ++k) run(k, 0);              // map building (untimed)
  run(k++, 1);                                     // sizes and page-locks the mirror (untimed)
  typedef std::chrono::steady_clock clk;
  const clk::time_point t0 = clk::now();
  for (int e = k + k_state; k < e; ++k) run(k, 1);
  const clk::time_point t1 = clk::now();
  for (int e = k + k_state; k < e; ++k) run(k, 2);
  const clk::time_point t1b = clk::now();
  run(k++, 0);
  const clk::time_point t2 = clk::now();
  for (int e = k + k_pose; k < e; ++k) run(k, 0);
  const clk::time_point t3 = clk::now();
  const int n = static_cast<int>(slam->GetStateVector().rows());
  std::printf("{\"n\": %d, \"state_steps\": %d, \"state_seconds\": %.6f, \"ref_seconds\": %.6f, \"pose_steps\": %d, \"pose_seconds\": %.6f, \"sink\": %.3e}\n", n,
              k_state, std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t1b - t1).count(), k_pose,
              std::chrono::duration<double>(t3 - t2).count(), sink);
  return 0;
}
