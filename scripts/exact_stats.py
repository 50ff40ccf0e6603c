"""How often, and for how many slots, the int8 SYRK defers to fp64 (per steady step of a C3 stream)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import drive_engine
from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM
from reflector_ekf_slam_b200.synth import make_stream
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
st = make_stream(cfg, steps)
e = ReflectorEKFSLAM(odom_model=st["model"], max_landmarks=st["N"], max_observations=st["m"], cov_update=2)
hist = []
for k in range(len(st["odom"])):
    drive_engine(e, st, k)
    raw = e.debug_copy("state", 20, np.int32)
    hist.append((int(raw[18]), int(raw[19]), int(raw[15])))   # exact_update, exact_slots, r
h = np.array(hist[st["n_build"]:])
print(f"{cfg}: {len(h)} steady steps; frames routed to full fp64 SYRK: {int(h[:,0].sum())}; frames with flagged slots: {int((h[:,1]>0).sum())};"
      f" flagged slots/frame mean {h[:,1].mean():.2f} max {h[:,1].max()}")
print("first 12 steps (exact_update, slots):", [tuple(x[:2]) for x in h[:12].tolist()])
idx = np.nonzero(h[:,1] > 0)[0]
print("steps with flagged slots:", idx[:40].tolist(), "...")
