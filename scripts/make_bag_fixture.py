"""Builds tests/golden/bag_stream.npz from the reference's shipped dataset (config C1).

Runs HERE only (reads /root/reference/dataset/*.bag, which does not exist on the GPU box); the resulting fixture is
committed.  Contents = exactly what the reference's node would feed its EKF for this bag, produced by the replay front-end
(reflector_ekf_slam_b200/replay: rosbag-v2 reader, LaserReflectorDetect + PoseExtrapolator restatement — pinned to the
reference's own detector by tests/test_replay.py) with the launch-file parameters (launch/slam.launch:24-27, ros_node.cc:240-283).
The node's call pattern (ros_node.cc:421-441, :627-660): the first scan only constructs the EKF with init_time = its stamp
(odometry before it is ignored because slam_ is null), later scans go through the detector into HandleObservationMessage
(empty frames included), every odometry message into HandleOdometryMessage."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from reflector_ekf_slam_b200.replay.__main__ import replay
    bags = glob.glob("/root/reference/dataset/*.bag")
    assert bags, "reference dataset not found"
    kind, time, odom_v, obs_start, obs_xy, box = [], [], [], [], [], {}

    def on_odom(t, vx, vy, wz):
        kind.append(0); time.append(t); odom_v.append((vx, vy, wz)); obs_start.append(len(obs_xy))

    def on_obs(t, xy):
        kind.append(1); time.append(t); odom_v.append((0.0, 0.0, 0.0)); obs_start.append(len(obs_xy))
        obs_xy.extend(xy.tolist())

    stats = replay(bags[0], on_odom, on_obs, lambda t0: box.setdefault("t0", t0))
    obs_start.append(len(obs_xy))
    out = os.path.join(ROOT, "tests", "golden", "bag_stream.npz")
    np.savez_compressed(out, init_time=np.float64(box["t0"]), kind=np.array(kind, np.int8), time=np.array(time, np.float64),
                        odom_v=np.array(odom_v, np.float64), obs_start=np.array(obs_start, np.int32),
                        obs_xy=np.array(obs_xy, np.float32).reshape(-1, 2))
    print(f"{os.path.basename(bags[0])}: {stats} -> {out}")


if __name__ == "__main__":
    main()
