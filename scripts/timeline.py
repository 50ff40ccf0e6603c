"""Kernel-start timeline of a pipelined replay (run with REKF_TIMELINE=1): which group's kernel starts when."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from reflector_ekf_slam_b200.engine import EKFBatch
from reflector_ekf_slam_b200.synth import make_stream
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T = 30
sts = [make_stream("C3", T, session=s) for s in range(S)]
b = EKFBatch(S, max_landmarks=1024, max_observations=100, cov_update=2, use_graphs=1, pipeline_groups=G)
nb = sts[0]["n_build"]
for k in range(nb):
    b.handle_odometry(np.stack([st["odom"][k] for st in sts]))
    b.handle_observation(np.array([st["obs_time"][k] for st in sts]), np.stack([st["obs_xy"][k] for st in sts]),
                         np.array([st["obs_count"][k] for st in sts]))
b.sync()
dev = torch.device("cuda:0")
d_odom = torch.tensor(np.stack([st["odom"][nb:] for st in sts]), device=dev)
d_time = torch.tensor(np.stack([st["obs_time"][nb:] for st in sts]), device=dev)
d_xy = torch.tensor(np.stack([st["obs_xy"][nb:] for st in sts]), device=dev)
torch.cuda.synchronize()
b.replay_device(d_odom.data_ptr(), d_time.data_ptr(), d_xy.data_ptr(), T, 100, None)
b.sync()
raw = b.debug_copy("tlog", 1 << 16, dtype=np.uint64)
n = int(raw[0])
e = raw[1:1 + n]
t = (e >> np.uint64(12)).astype(np.float64) / 1e3
kid = ((e >> np.uint64(8)) & np.uint64(15)).astype(int)
s0 = (e & np.uint64(255)).astype(int)
names = ["odo", "front", "innov", "chol", "solve", "syrkf64", "syrk", "augment", "gather_y", "chol_end", "lastflag", "solve_end", "?", "?", "?", "?"]
order = np.argsort(t)
t, kid, s0 = t[order], kid[order], s0[order]
fronts = np.nonzero((kid == 1) & (s0 == 0))[0]        # group 0's front kernels: one per step
sel = slice(fronts[-4], fronts[-1])
t0 = t[sel][0]
prev = {}
for tt, k, g in zip(t[sel], kid[sel], s0[sel]):
    d = tt - prev.get(g, tt)
    prev[g] = tt
    print(f"{tt - t0:9.1f} us  group@{g:<2d} {names[k]:8s} (+{d:6.1f} since this group's previous mark)")
print("steps/s over the last 20 steps:", 20 * S / ((t[fronts[-1]] - t[fronts[-21]]) * 1e-6))
