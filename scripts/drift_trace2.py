"""Like drift_trace.py, but the tensor-mode engine takes over from the fp64 engine's state after `handover`
steady steps (isolates ongoing per-step error from the first high-cancellation update)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import drive_engine, rel_fro  # noqa: E402
from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM  # noqa: E402
from reflector_ekf_slam_b200.synth import make_stream  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
handover = int(sys.argv[4]) if len(sys.argv) > 4 else 1
every = int(sys.argv[5]) if len(sys.argv) > 5 else 1
st = make_stream(cfg, steps)
kw = dict(odom_model=st["model"], max_landmarks=st["N"], max_observations=st["m"])
a = ReflectorEKFSLAM(cov_update=1, **kw)
b = ReflectorEKFSLAM(cov_update=mode, **kw)
for k in range(len(st["odom"])):
    j = k - st["n_build"] + 1
    drive_engine(a, st, k)
    if j < handover:
        continue
    if j == handover:
        t, mu, sig = a.GetState()
        b.set_state(t, st["odom"][k][1:4], mu, sig)
        continue
    drive_engine(b, st, k)
    if (j - handover) % every == 0 or k == len(st["odom"]) - 1:
        Sa, Sb = a.GetCoviarance(), b.GetCoviarance()
        d = np.sqrt(np.abs(np.diag(Sa)))
        E = np.abs(Sb - Sa) / np.outer(d, d)
        print(f"step {j:3d}: relFro {rel_fro(Sb, Sa):.2e}  max corr-normalised err {E.max():.2e}  "
              f"|dmu| {np.abs(a.GetStateVector() - b.GetStateVector()).max():.2e}  min eig {np.linalg.eigvalsh(Sb)[0]:.2e}", flush=True)
