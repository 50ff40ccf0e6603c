"""Per-session total cycles of k_cholesky_smem in a batch of S sessions (build with REKF_NVCC_EXTRA=-DREKF_CHOL_TIMING)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflector_ekf_slam_b200.engine import EKFBatch
from reflector_ekf_slam_b200.synth import make_stream
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sts = [make_stream("C3", 3, session=s) for s in range(S)]
b = EKFBatch(S, max_landmarks=1024, max_observations=100, cov_update=2)
for k in range(len(sts[0]["odom"])):
    b.handle_odometry(np.stack([st["odom"][k] for st in sts]))
    b.handle_observation(np.array([st["obs_time"][k] for st in sts]), np.stack([st["obs_xy"][k] for st in sts]),
                         np.array([st["obs_count"][k] for st in sts]))
b.sync()
nb = 7
for s in range(S):
    t = b.debug_copy("qd", 72, s=s)
    d = np.diff(t[:3 + 3 * nb])
    ph = d[1:1 + 3 * nb].reshape(nb, 3).sum(0).astype(int)
    if s == 0:
        print("  per block [p1, p2, p3]:", d[1:1 + 3 * nb].reshape(nb, 3).astype(int).tolist())
    print(f"session {s}: start {int(t[0]) % 10**9:10d} load {int(d[0]):6d} phases {ph.tolist()} publish {int(d[-1]):6d} total {int(t[2 + 3 * nb] - t[0]):7d}")
