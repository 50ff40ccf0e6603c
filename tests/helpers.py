"""Shared drivers for the parity tests: run one synthetic stream through a filter, step by step."""
import numpy as np

MU_TOL_M = 1e-4          # BASELINE.json north_star: pose / landmark means within 1e-4 m
SIGMA_REL_FRO = 1e-5     # covariance within 1e-5 relative Frobenius


def rel_fro(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def drive_oracle(oracle, stream, k):
    o = stream["odom"][k]
    c = int(stream["obs_count"][k])
    oracle.HandleOdometryMessage(o[0], o[1], o[2], o[3])
    oracle.HandleObservationMessage(stream["obs_time"][k], stream["obs_xy"][k, :c])


def drive_engine(ekf, stream, k):
    from reflector_ekf_slam_b200.engine import Observation, OdometryData
    o = stream["odom"][k]
    c = int(stream["obs_count"][k])
    ekf.HandleOdometryMessage(OdometryData(o[0], o[1], o[2], o[3]))
    ekf.HandleObservationMessage(Observation(stream["obs_time"][k], stream["obs_xy"][k, :c]))


def compare_state(ekf, oracle, check_sigma=True, tag=""):
    mu_g, mu_o = ekf.GetStateVector(), oracle.GetStateVector()
    assert mu_g.shape == mu_o.shape, f"{tag}: state size {mu_g.shape} vs oracle {mu_o.shape}"
    dmu = float(np.abs(mu_g - mu_o).max())
    assert dmu < MU_TOL_M, f"{tag}: |mu - oracle|_max = {dmu:.3e} m"
    ds = 0.0
    if check_sigma:
        ds = rel_fro(ekf.GetCoviarance(), oracle.GetCoviarance())
        assert ds < SIGMA_REL_FRO, f"{tag}: Sigma relative Frobenius error {ds:.3e}"
    return dmu, ds


def compare_matches(ekf, oracle, tag=""):
    g, o = ekf.match_result(), oracle.match_result()
    for a, b, name in zip(g, o, ("state_obs_match_ids", "map_obs_match_ids", "new_ids")):
        assert np.array_equal(a, b), f"{tag}: {name} differ\nengine {a.tolist()}\noracle {b.tolist()}"
