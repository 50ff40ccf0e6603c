"""GPU parity against the REFERENCE'S OWN CODE (run with -m gpu on a B200): the CUDA engine, through the C ABI, against
ekf::ReflectorEKFSLAM / ekf::ReflectorEKFSLAMGPS compiled unmodified from /root/reference into oracle/_ref/
(librekf_ref.so travels to the GPU box with the snapshot; nothing here reads /root/reference at run time).

Same seeded inputs, every step: association lists identical, means within 1e-4 m, covariance within 1e-5 relative
Frobenius (BASELINE.json).  The restated oracle does not appear in this file."""
import os

import numpy as np
import pytest

from helpers import compare_matches, compare_state, drive_engine, drive_oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "librekf_ref.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/librekf_ref.so not built (make -C oracle ref where /root/reference exists)")
COV = {"f64": 1, "i8": 2}


def _pair(st, cov, gps=False, **kw):
    from oracle import pyoracle
    from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM
    ekf = ReflectorEKFSLAM(odom_model=st["model"], max_landmarks=max(st["N"], 8), max_observations=max(st["m"], 8), cov_update=COV[cov], **kw)
    ref = pyoracle.Reference(odom_model=st["model"], gps=gps, fast=True, **{k: v for k, v in kw.items() if k in ("map_path", "init_time")})
    return ekf, ref


@needs_ref
@pytest.mark.parametrize("cov", ["f64", "i8"])
@pytest.mark.parametrize("cfg,steps", [("T0", 40), ("T1", 30)])
def test_small_streams_every_step_vs_reference(engine_lib, cfg, steps, cov):
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream(cfg, steps)
    ekf, ref = _pair(st, cov)
    worst = (0.0, 0.0)
    for k in range(len(st["odom"])):
        drive_engine(ekf, st, k)
        drive_oracle(ref, st, k)
        compare_matches(ekf, ref, f"{cfg} step {k}")
        d = compare_state(ekf, ref, tag=f"{cfg}/{cov} step {k} vs reference")
        worst = (max(worst[0], d[0]), max(worst[1], d[1]))
    assert ekf.error_flags() == 0 and ekf.GetLatestTime() == ref.GetLatestTime()
    print(f"{cfg}/{cov} vs reference: worst |dmu| {worst[0]:.2e} m, worst rel-Fro {worst[1]:.2e}")


@needs_ref
def test_c2_stream_vs_reference(engine_lib):
    """Config C2 (N=256, m=50, n=515): map building through both augmentation paths + 12 steady steps."""
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("C2", 12)
    ekf, ref = _pair(st, "i8")
    for k in range(len(st["odom"])):
        drive_engine(ekf, st, k)
        drive_oracle(ref, st, k)
        compare_matches(ekf, ref, f"C2 step {k}")
        compare_state(ekf, ref, check_sigma=(k % 3 == 0 or k == len(st["odom"]) - 1), tag=f"C2 step {k} vs reference")
    assert ekf.dim() == 515 and ekf.error_flags() == 0


@needs_ref
@pytest.mark.parametrize("cov", ["i8", "f64"])
def test_c3_steps_vs_reference(engine_lib, cov):
    """Config C3 (N=1024, m=100, n=2051), the headline size: the engine builds the map through its own augmentation, the
    reference starts from that snapshot (its as-written map building would take minutes), then 3 steps of the reference's
    dense algebra (seconds each) against the engine."""
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("C3", 3)
    ekf, ref = _pair(st, cov)
    nb = st["n_build"]
    for k in range(nb):
        drive_engine(ekf, st, k)
    t, mu, S = ekf.GetState()
    assert mu.size == 2051
    ref.set_state(t, st["odom"][nb - 1][1:4], mu, S)
    for k in range(nb, nb + 3):
        drive_engine(ekf, st, k)
        drive_oracle(ref, st, k)
        compare_matches(ekf, ref, f"C3 step {k}")
        dmu, ds = compare_state(ekf, ref, tag=f"C3/{cov} step {k} vs reference")
    assert len(ekf.match_result()[0]) == 100 and ekf.error_flags() == 0
    print(f"C3/{cov} vs reference after 3 steps: |dmu| {dmu:.2e} m, rel-Fro {ds:.2e}")


@needs_ref
def test_gps_class_vs_reference(engine_lib):
    """ekf::ReflectorEKFSLAMGPS (reflector_ekf_slam_gps.cc:305-340): three pose rows on top of the reflector rows."""
    from reflector_ekf_slam_b200.engine import Observation, OdometryData
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("T0", 20)
    ekf, ref = _pair(st, "f64", gps=True)
    rng = np.random.default_rng(5)
    for k in range(len(st["odom"])):
        o, c = st["odom"][k], int(st["obs_count"][k])
        gps = st["true_pose"][k] + rng.normal(0, [0.03, 0.03, 0.01])
        ekf.HandleOdometryMessage(OdometryData(*o))
        ref.HandleOdometryMessage(*o)
        ekf.HandleObservationMessage(Observation(st["obs_time"][k], st["obs_xy"][k, :c], gps_pose=gps))
        ref.HandleObservationMessage(st["obs_time"][k], st["obs_xy"][k, :c], gps_pose=gps)
        compare_matches(ekf, ref, f"gps step {k}")
        compare_state(ekf, ref, tag=f"gps step {k} vs reference")


@needs_ref
def test_map_loader_reference_mode_and_beacon_localisation_vs_reference(engine_lib, tmp_path):
    """rekf_load_map_txt with REKF_MAP_LOADER_REFERENCE reads what the reference's LoadMapFromTxtFile reads (:87-91: the
    covariances come from the POSITIONS line; entries the reference reads past the end of that line — undefined there —
    are 0.0 here), and the beacon-map branch (:401-425, :279-303) follows the reference step by step."""
    from oracle import pyoracle
    from reflector_ekf_slam_b200._abi import REKF_MAP_LOADER_REFERENCE
    from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("T0", 30)
    b = ReflectorEKFSLAM(max_landmarks=16, max_observations=8, cov_update=1)
    for k in range(st["n_build"] + 4):
        drive_engine(b, st, k)
    xy, cov = b.landmarks()
    path = str(tmp_path / "map.txt")
    with open(path, "w") as f:                              # well-formed two-line file (no leading comma)
        f.write(",".join("%g,%g" % (x, y) for x, y in xy) + "\n")
        f.write(",".join("%g,%g,%g,%g" % tuple(c.reshape(-1)) for c in cov) + "\n")
    eng = ReflectorEKFSLAM(max_landmarks=16, max_observations=8, cov_update=1, map_path=path, map_loader=REKF_MAP_LOADER_REFERENCE)
    ref = pyoracle.Reference(map_path=path)
    exy, ecov = eng.GetGlobalMap()
    rxy, rcov = ref.GetGlobalMap()
    assert exy.shape == rxy.shape == (16, 2) and np.array_equal(exy, rxy)
    assert np.array_equal(ecov[:8], rcov[:8])               # 32 numbers on line 1 = the first 8 "covariances" the reference reads in bounds
    assert np.all(ecov[8:] == 0.0)
    # the fixed loader reads line 2
    fixed = ReflectorEKFSLAM(max_landmarks=16, max_observations=8, cov_update=1, map_path=path)
    assert np.allclose(fixed.GetGlobalMap()[1], cov, rtol=2e-5, atol=1e-12)
    # beacon localisation against the reference with the same injected map
    small = np.tile(np.eye(2) * 1e-3, (16, 1, 1))
    eng2 = ReflectorEKFSLAM(max_landmarks=16, max_observations=8, cov_update=1)
    ref2 = pyoracle.Reference()
    eng2.set_map(rxy, small)
    ref2.set_map(rxy, small)
    n_map = 0
    for k in range(len(st["odom"])):
        drive_engine(eng2, st, k)
        drive_oracle(ref2, st, k)
        compare_matches(eng2, ref2, f"beacon step {k}")
        compare_state(eng2, ref2, tag=f"beacon step {k} vs reference")
        n_map += len(eng2.match_result()[1])
    assert n_map > 20


def test_get_state_single_sync_and_counters(engine_lib):
    """rekf_get_state == the separate getters; a wrong expected dimension reports the real one; rekf_get_counters counts updates."""
    from reflector_ekf_slam_b200.engine import RekfError, ReflectorEKFSLAM
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("T1", 6)
    ekf = ReflectorEKFSLAM(odom_model=st["model"], max_landmarks=60, max_observations=12)
    for k in range(len(st["odom"])):
        drive_engine(ekf, st, k)
    t, mu, sig, flags = ekf.state()
    assert flags == 0 and t == ekf.GetLatestTime()
    assert np.array_equal(mu, ekf.GetStateVector()) and np.array_equal(sig, ekf.GetCoviarance())
    assert np.array_equal(sig, sig.T)
    with pytest.raises(RekfError) as e:
        ekf.state(n_expect=mu.size - 2)
    assert e.value.code == -3
    c = ekf.counters()
    assert c["updates"] == 6 and c["exact_frames"] + c["updates"] >= 6
