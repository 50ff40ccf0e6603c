"""ROS-free replay front-end: rosbag-v2 reader -> LaserReflectorDetect / PoseExtrapolator (CPU restatement
of the reference's detector, reference src/reflector_detect/laser/) -> the engine's C ABI, with the node's call pattern
(reference src/ros_node.cc:421-441, :627-660), and the reference's landmark-map text file at the end.

    python -m reflector_ekf_slam_b200.replay <file.bag> [--out <filebase>]

The EKF itself runs on the GPU through librekf_b200.so; nothing here computes a filter step on the CPU."""
from .bag import read_bag, parse_odometry, parse_scan  # noqa: F401
from .laser_detect import DetectOptions, LaserReflectorDetect  # noqa: F401
from .pose_extrapolator import OdometrySample, PoseExtrapolator  # noqa: F401
