"""Per-CTA ramp / tail of the persistent SYRK (build with REKF_NVCC_EXTRA=-DREKF_SYRK_TIMING)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflector_ekf_slam_b200.engine import EKFBatch
from reflector_ekf_slam_b200.synth import make_stream
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sts = [make_stream("C3", 4, session=s) for s in range(S)]
b = EKFBatch(S, max_landmarks=1024, max_observations=100, cov_update=2)
for k in range(len(sts[0]["odom"])):
    b.handle_odometry(np.stack([st["odom"][k] for st in sts]))
    b.handle_observation(np.array([st["obs_time"][k] for st in sts]), np.stack([st["obs_xy"][k] for st in sts]),
                         np.array([st["obs_count"][k] for st in sts]))
b.sync()
raw = np.concatenate([b.debug_copy("innov", 204, s=s) for s in range(S)])
n = min(148, raw.size // 6)
t = raw[: n * 6].reshape(n, 6)
t0 = t[:, 0].min()
us = (t[:, :5] - t0) / 1e3
for name, col in (("start", 0), ("setup done", 1), ("first tile done", 2), ("last tile done", 3), ("exit", 4)):
    print(f"{name:16s} min {us[:, col].min():8.1f}  median {np.median(us[:, col]):8.1f}  max {us[:, col].max():8.1f} us")
print("tiles per CTA: min", int(t[:, 5].min()), "median", int(np.median(t[:, 5])), "max", int(t[:, 5].max()), " CTAs", n)
