"""Per-phase cycle counts of k_cholesky_smem (build with REKF_NVCC_EXTRA=-DREKF_CHOL_TIMING)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import drive_engine
from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM
from reflector_ekf_slam_b200.synth import make_stream
st = make_stream("C3", 3)
e = ReflectorEKFSLAM(max_landmarks=1024, max_observations=100, cov_update=2)
for k in range(len(st["odom"])):
    drive_engine(e, st, k)
nb = 7
t = e.debug_copy("qd", 72)
print("load", int(t[1] - t[0]))
names = ["1(diag || prev trailing)", "2(rows)", "3(next column)"]
tot = np.zeros(3)
for b in range(nb):
    s = t[1 + 3 * b: 1 + 3 * b + 4]
    d = np.diff(s)
    tot += d
    print(f"block {b}: " + "  ".join(f"{n} {int(x):6d}" for n, x in zip(names, d)))
print("publish", int(t[2 + 3 * nb] - t[1 + 3 * nb]))
print("totals:", dict(zip(names, tot.astype(int))), "sum", int(t[2 + 3 * nb] - t[0]), "cycles")
print("per-warp finish - start:", [int(x - t[0]) for x in t[64:72]])
