"""Per-step divergence of a tensor-core covariance mode from the fp64 SIMT mode (both on the GPU)."""
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
st = make_stream(cfg, steps)
kw = dict(odom_model=st["model"], max_landmarks=st["N"], max_observations=st["m"])
a = ReflectorEKFSLAM(cov_update=1, **kw)
b = ReflectorEKFSLAM(cov_update=mode, **kw)
for k in range(len(st["odom"])):
    drive_engine(a, st, k)
    drive_engine(b, st, k)
    j = k - st["n_build"] + 1
    if j >= -2:
        Sa, Sb = a.GetCoviarance(), b.GetCoviarance()
        d = np.sqrt(np.abs(np.diag(Sa)))
        E = np.abs(Sb - Sa) / np.outer(d, d)
        i0, j0 = np.unravel_index(np.argmax(E), E.shape)
        pose = np.abs(Sb[:3, :3] - Sa[:3, :3]).max() / np.abs(Sa[:3, :3]).max()
        ev = np.linalg.eigvalsh(Sb)[0]
        print(f"step {j:3d}: relFro {rel_fro(Sb, Sa):.2e}  max corr-normalised err {E.max():.2e} at ({i0},{j0})  pose-block rel {pose:.2e}"
              f"  |dmu| {np.abs(a.GetStateVector() - b.GetStateVector()).max():.2e}  min eig {ev:.2e}  matches {len(b.match_result()[0])}", flush=True)
