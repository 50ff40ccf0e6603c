"""Per-phase cycle counts of k_solve_w3, CTA 0 (build with REKF_NVCC_EXTRA=-DREKF_SOLVE_TIMING)."""
import os, sys
os.environ.setdefault("REKF_SOLVE_LL", "0")   # the instrumented kernel is the previous-generation k_solve_w3
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
t = e.debug_copy("innov", 2 + 2 * 7 + 3)
d = np.diff(t).astype(int)
print("gather", d[0])
for b in range(7):
    print(f"block {b}: wait/stageX {d[1 + 2 * b]:6d}   W_J {d[2 + 2 * b] if 2 + 2 * b < len(d) else 0:6d} (incl. previous block's trailing chunks in 'wait')")
print("rest:", d[15:].tolist(), "total", int(t[-1] - t[0]))
