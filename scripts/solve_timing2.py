"""All cycle stamps of k_solve_w3, CTA 0 of session 0, with S sessions running side by side
(build with REKF_NVCC_EXTRA="-DREKF_SOLVE_TIMING -DREKF_SOLVE_TIMING2").  usage: solve_timing2.py [S]"""
import os, sys
os.environ.setdefault("REKF_SOLVE_LL", "0")   # the instrumented kernel is the previous-generation k_solve_w3
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflector_ekf_slam_b200.engine import EKFBatch
from reflector_ekf_slam_b200.synth import make_stream
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sts = [make_stream("C3", 3 + s) for s in range(S)]
e = EKFBatch(S, max_landmarks=1024, max_observations=100, cov_update=2)
T = len(sts[0]["odom"])
for k in range(T):
    e.handle_odometry(np.stack([st["odom"][k] for st in sts]))
    m = max(int(st["obs_count"][k]) for st in sts)
    xy = np.zeros((S, max(m, 1), 2), np.float32)
    for s, st in enumerate(sts):
        xy[s, :int(st["obs_count"][k])] = st["obs_xy"][k, :int(st["obs_count"][k])]
    e.handle_observation(np.array([st["obs_time"][k] for st in sts]), xy, np.array([int(st["obs_count"][k]) for st in sts], np.int32))
e.sync()
print("stamps: start | desc staged | gather loads done | sync | (wait, W_J) x7 | solve end | mu | W64 | Wq digits | end")
for s in sorted({0, S - 1}):
    t = e.debug_copy("innov", 2 + 2 + 2 * 7 + 3 + 2, s=s)
    print(s, np.diff(t).astype(int).tolist(), "total", int(t[-1] - t[0]))
