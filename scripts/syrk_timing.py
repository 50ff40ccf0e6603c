"""Where every role of the persistent SYRK waits (build with REKF_NVCC_EXTRA=-DREKF_SYRK_TIMING).
usage: python scripts/syrk_timing.py [sessions=4]   -> per-role mean cycles per CTA spent in each wait"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reflector_ekf_slam_b200.engine import EKFBatch
from reflector_ekf_slam_b200.synth import make_stream
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sts = [make_stream("C3", 4, session=s) for s in range(S)]
b = EKFBatch(S, max_landmarks=1024, max_observations=100, cov_update=2)
for k in range(len(sts[0]["odom"])):
    b.handle_odometry(np.stack([st["odom"][k] for st in sts]))
    b.handle_observation(np.array([st["obs_time"][k] for st in sts]), np.stack([st["obs_xy"][k] for st in sts]),
                         np.array([st["obs_count"][k] for st in sts]))
b.sync()
raw = b.debug_copy("sbuf", 148 * 24, s=0).reshape(148, 24)
t = raw[raw[:, 20] > 0]
print(f"{len(t)} CTAs with tiles; tiles per CTA min/mean/max {t[:,20].min():.0f}/{t[:,20].mean():.1f}/{t[:,20].max():.0f}")
names = {
    "producer": [(0, "wait tile-ring slot"), (1, "wait A chunk free"), (2, "wait B stage free"), (3, "TOTAL"), (4, "tiles published")],
    "mma": [(5, "wait tile ring"), (6, "wait acc set free"), (7, "wait A chunk"), (8, "wait B stage"), (9, "TOTAL")],
    "reduce": [(10, "wait tile ring"), (11, "wait box written"), (12, "wait TMA read-out"), (13, "final wait"), (14, "TOTAL")],
    "epilogue": [(15, "wait tile ring"), (16, "wait scales"), (17, "wait accumulators"), (18, "wait box slot free"), (19, "TOTAL")],
}
for role, cols in names.items():
    print(role)
    for c, nm in cols:
        print(f"   {nm:24s} mean {t[:, c].mean():10.0f}  min {t[:, c].min():10.0f}  max {t[:, c].max():10.0f}  cycles   ({t[:, c].mean() / 1965:7.2f} us)")
