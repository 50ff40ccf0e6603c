"""One-step error of each covariance-update mode against the fp64 SIMT mode, from the same state.
Usage: python scripts/syrk_error.py [config] [steady_steps_before]"""
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
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 3
st = make_stream(cfg, pre + 2)
kw = dict(odom_model=st["model"], max_landmarks=st["N"], max_observations=st["m"])
base = ReflectorEKFSLAM(cov_update=1, **kw)
for k in range(st["n_build"] + pre):
    drive_engine(base, st, k)
t, mu, sig = base.GetState()
vt = st["odom"][st["n_build"] + pre - 1][1:4]
k = st["n_build"] + pre
res = {}
for name, mode in (("f64", 1), ("tf32x3", 0), ("i8x4", 2)):
    e = ReflectorEKFSLAM(cov_update=mode, **kw)
    e.set_state(t, vt, mu, sig)
    drive_engine(e, st, k)
    res[name] = (e.GetStateVector(), e.GetCoviarance())
S0 = res["f64"][1]
dS = S0 - sig
print(f"{cfg}: n={mu.size}  |dSigma|_F/|Sigma|_F = {np.linalg.norm(dS)/np.linalg.norm(sig):.3e} (includes predict noise)")
for name in ("tf32x3", "i8x4"):
    S = res[name][1]
    print(f"  {name}: relFro vs f64 = {rel_fro(S, S0):.3e}   max|dS| = {np.abs(S-S0).max():.3e}   |dmu| = {np.abs(res[name][0]-res['f64'][0]).max():.3e}"
          f"   asym = {np.abs(S-S.T).max():.1e}")
    D = np.abs(S - S0)
    i, j = np.unravel_index(np.argmax(D), D.shape)
    print(f"     worst element ({i},{j}): f64 {S0[i,j]:.6e} vs {S[i,j]:.6e}; prior {sig[i,j]:.6e}")
