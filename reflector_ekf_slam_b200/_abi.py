"""ctypes mirror of include/rekf.h (struct rekf_options and the enums).

Used by the engine binding (engine.py).  The struct is plain data; the test-side CPU checker reuses it.
"""
import ctypes as C

REKF_ODOM_DIFF = 0
REKF_ODOM_OMNI = 1

REKF_COV_TCGEN05_TF32X3 = 0
REKF_COV_SIMT_F64 = 1
REKF_COV_TCGEN05_I8X4 = 2

REKF_MAP_LOADER_FIXED = 0
REKF_MAP_LOADER_REFERENCE = 1

REKF_OK = 0
REKF_ERR_BAD_ARGUMENT = -1
REKF_ERR_CUDA = -2
REKF_ERR_CAPACITY = -3
REKF_ERR_NOT_SPD = -4
REKF_ERR_IO = -5
REKF_ERR_NO_DEVICE = -6
REKF_ERR_UNSUPPORTED = -7


class RekfOptions(C.Structure):
    """struct rekf_options (include/rekf.h) = ekf::EKFOptions (ekf_slam_interface.h:28-41) + engine fields."""

    _fields_ = [
        ("use_imu", C.c_int),
        ("init_time", C.c_double),
        ("init_pose", C.c_double * 3),
        ("map_path", C.c_char_p),
        ("odom_model", C.c_int),
        ("linear_velocity_cov", C.c_double),
        ("angular_velocity_cov", C.c_double),
        ("observation_cov", C.c_double),
        ("max_landmarks", C.c_int),
        ("max_observations", C.c_int),
        ("max_map_landmarks", C.c_int),
        ("device", C.c_int),
        ("cov_update", C.c_int),
        ("map_loader", C.c_int),
        ("stream", C.c_void_p),
        ("use_graphs", C.c_int),
        ("pipeline_groups", C.c_int),
        ("syrk_reserve_sms", C.c_int),
    ]


def make_options(
    init_time=0.0,
    init_pose=(0.0, 0.0, 0.0),
    map_path=None,
    odom_model=REKF_ODOM_DIFF,
    linear_velocity_cov=0.05 * 0.05,
    angular_velocity_cov=0.08 * 0.08,
    observation_cov=0.05 * 0.05,
    max_landmarks=0,
    max_observations=0,
    max_map_landmarks=0,
    device=0,
    cov_update=REKF_COV_TCGEN05_I8X4,
    map_loader=REKF_MAP_LOADER_FIXED,
    stream=None,
    use_graphs=0,
    use_imu=False,
    pipeline_groups=0,
    syrk_reserve_sms=0,
):
    """Options with the reference's launch-file defaults (launch/slam.launch:21-23: sigma_v 0.05,
    sigma_w 0.08, sigma_z 0.05; squared the way ros_node.cc:207-237 squares them)."""
    o = RekfOptions()
    o.use_imu = int(bool(use_imu))
    o.init_time = float(init_time)
    o.init_pose[0], o.init_pose[1], o.init_pose[2] = (float(v) for v in init_pose)
    o.map_path = map_path.encode() if isinstance(map_path, str) else map_path
    o.odom_model = int(odom_model)
    o.linear_velocity_cov = float(linear_velocity_cov)
    o.angular_velocity_cov = float(angular_velocity_cov)
    o.observation_cov = float(observation_cov)
    o.max_landmarks = int(max_landmarks)
    o.max_observations = int(max_observations)
    o.max_map_landmarks = int(max_map_landmarks)
    o.device = int(device)
    o.cov_update = int(cov_update)
    o.map_loader = int(map_loader)
    o.stream = stream
    o.use_graphs = int(use_graphs)
    o.pipeline_groups = int(pipeline_groups)
    o.syrk_reserve_sms = int(syrk_reserve_sms)
    return o
