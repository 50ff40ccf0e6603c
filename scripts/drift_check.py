"""Long-horizon drift of the tcgen05 (tf32x3) covariance path against the fp64 CPU oracle.
Usage: python scripts/drift_check.py [config] [steps] [every]   (run on a GPU box)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import drive_engine, drive_oracle, rel_fro  # noqa: E402
from oracle.pyoracle import STRUCTURED, Oracle  # noqa: E402
from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM  # noqa: E402
from reflector_ekf_slam_b200.synth import make_stream  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
every = int(sys.argv[3]) if len(sys.argv) > 3 else 50
st = make_stream(cfg, steps)
engines = {name: ReflectorEKFSLAM(odom_model=st["model"], max_landmarks=st["N"], max_observations=st["m"], cov_update=mode)
           for name, mode in (("tf32x3", 0), ("f64", 1), ("i8x4", 2))}
orc = Oracle(algebra=STRUCTURED, odom_model=st["model"], native=True)
t0 = time.time()
for k in range(len(st["odom"])):
    for e in engines.values():
        drive_engine(e, st, k)
    drive_oracle(orc, st, k)
    j = k - st["n_build"] + 1
    if j > 0 and (j % every == 0 or k == len(st["odom"]) - 1):
        mu_o, S_o = orc.GetStateVector(), orc.GetCoviarance()
        line = f"{cfg} steady step {j:4d}:"
        for name, e in engines.items():
            line += f"  {name}: |dmu| {np.abs(e.GetStateVector() - mu_o).max():.2e} m  relFro {rel_fro(e.GetCoviarance(), S_o):.2e}"
        print(line + f"   [{time.time() - t0:.0f}s]", flush=True)
