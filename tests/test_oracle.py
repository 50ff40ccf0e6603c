"""CPU tests of the oracle itself (no GPU): the C restatement against hand-derived known answers, its own
two algebra modes, and the independently written numpy EKF (oracle/numpy_ekf.py).

The reference has no tests or golden vectors for this path — these are what pin the oracle."""
import math
import os

import numpy as np
import pytest

from oracle.numpy_ekf import NumpyEKF
from oracle.pyoracle import AS_WRITTEN, STRUCTURED, Oracle, load
from reflector_ekf_slam_b200.synth import DIFF, OMNI, make_stream

from helpers import rel_fro


def test_dgemm_matches_numpy():
    import ctypes as C
    lib = load()
    rng = np.random.default_rng(1)
    for (m, n, k, ta, tb) in [(7, 5, 3, 0, 0), (130, 67, 300, 0, 1), (33, 129, 257, 1, 0), (64, 64, 64, 1, 1)]:
        A = np.asfortranarray(rng.normal(size=(k, m) if ta else (m, k)))
        B = np.asfortranarray(rng.normal(size=(n, k) if tb else (k, n)))
        Cm = np.zeros((m, n), order="F")
        lib.oracle_dgemm(ta, tb, m, n, k, A.ctypes.data_as(C.c_void_p), A.shape[0], B.ctypes.data_as(C.c_void_p), B.shape[0],
                         Cm.ctypes.data_as(C.c_void_p), m)
        ref = (A.T if ta else A) @ (B.T if tb else B)
        assert np.abs(Cm - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


def test_lu_inverse():
    import ctypes as C
    lib = load()
    rng = np.random.default_rng(2)
    for n in (1, 2, 6, 57, 200):
        X = rng.normal(size=(n, n))
        A = np.asfortranarray(X @ X.T + n * np.eye(n))
        A[0, :], A[-1, :] = A[-1, :].copy(), A[0, :].copy()      # force pivoting
        inv = A.copy(order="F")
        assert lib.oracle_lu_inverse(n, inv.ctypes.data_as(C.c_void_p), n) == 0
        assert np.abs(inv @ A - np.eye(n)).max() < 1e-10


@pytest.mark.parametrize("algebra", [AS_WRITTEN, STRUCTURED])
def test_known_answer_single_predict_from_zero_covariance(algebra):
    """Σ₀ = 0 ⇒ after one predict Σ = G_u·Qu·G_uᵀ exactly (reflector_ekf_slam.cc:178)."""
    o = Oracle(algebra=algebra, init_pose=(1.0, 2.0, 0.3))
    v, w, dt = 0.7, 0.2, 0.1
    o.HandleOdometryMessage(dt, v, 0.0, w)
    a = 0.3 + w * dt / 2
    Gu = np.array([[dt * math.cos(a), -v * dt * dt * math.sin(a) / 2], [dt * math.sin(a), v * dt * dt * math.cos(a) / 2], [0, dt]])
    want = Gu @ np.diag([0.0025, 0.0064]) @ Gu.T
    assert np.allclose(o.GetCoviarance(), want, rtol=0, atol=1e-18)
    mu = o.GetStateVector()
    assert np.allclose(mu, [1.0 + v * dt * math.cos(a), 2.0 + v * dt * math.sin(a), 0.3 + w * dt], atol=1e-15)
    assert o.GetLatestTime() == dt


@pytest.mark.parametrize("algebra", [AS_WRITTEN, STRUCTURED])
def test_known_answer_first_landmark_initialisation(algebra):
    """One unmatched observation: Σ_mm = G_p·Σ_xx·G_pᵀ + Qt (:354), Σ_mx = G_p·Σ_x· (:355), mean through float32."""
    o = Oracle(algebra=algebra, init_pose=(0.5, -0.25, 0.1))
    o.HandleOdometryMessage(0.2, 0.4, 0.0, -0.1)
    P = o.GetCoviarance().copy()
    o.HandleObservationMessage(0.2, np.array([[2.0, 1.0]], np.float32))   # dt = 0 predict adds nothing
    mu, S = o.GetStateVector(), o.GetCoviarance()
    assert mu.size == 5
    th = mu[2]
    c, s = math.cos(th), math.sin(th)
    gx = np.float32(2.0 * c - 1.0 * s + mu[0])
    gy = np.float32(2.0 * s + 1.0 * c + mu[1])
    assert mu[3] == float(gx) and mu[4] == float(gy)
    Gp = np.array([[1, 0, -2.0 * s - 1.0 * c], [0, 1, 2.0 * c - 1.0 * s]])
    assert np.allclose(S[:3, :3], P, atol=1e-18)
    assert np.allclose(S[3:, :3], Gp @ P, atol=1e-17)
    assert np.allclose(S[:3, 3:], (Gp @ P).T, atol=1e-17)
    assert np.allclose(S[3:, 3:], Gp @ P @ Gp.T + 0.0025 * np.eye(2), atol=1e-17)
    sp, mp, nw = o.match_result()
    assert len(sp) == 0 and len(mp) == 0 and nw.tolist() == [0]


def test_known_answer_single_update_against_numpy_inverse():
    """One matched landmark: μ, Σ against K = ΣHᵀ(HΣHᵀ+Q)⁻¹ formed by hand with numpy.linalg.inv."""
    o = Oracle(algebra=AS_WRITTEN)
    o.HandleOdometryMessage(0.1, 0.5, 0.0, 0.05)
    o.HandleObservationMessage(0.1, np.array([[3.0, 0.5]], np.float32))          # creates landmark 0
    o.HandleOdometryMessage(0.2, 0.5, 0.0, 0.05)
    mu0, S0 = o.GetStateVector(), o.GetCoviarance()
    z = np.array([2.93, 0.52], np.float32)
    o.HandleObservationMessage(0.2, z.reshape(1, 2))
    sp, mp, nw = o.match_result()
    assert sp.tolist() == [[0, 0]] and len(nw) == 0
    c, s = math.cos(mu0[2]), math.sin(mu0[2])
    d = mu0[3:5] - mu0[:2]
    H = np.array([[-c, -s, -d[0] * s + d[1] * c, c, s], [s, -c, -d[0] * c - d[1] * s, -s, c]])
    zhat = np.array([d[0] * c + d[1] * s, -d[0] * s + d[1] * c])
    K = S0 @ H.T @ np.linalg.inv(H @ S0 @ H.T + 0.0025 * np.eye(2))
    mu1 = mu0 + K @ (z.astype(np.float64) - zhat)
    mu1[2] = math.atan2(math.sin(mu1[2]), math.cos(mu1[2]))
    assert np.allclose(o.GetStateVector(), mu1, atol=1e-14)
    assert np.allclose(o.GetCoviarance(), S0 - K @ H @ S0, atol=1e-16)


@pytest.mark.parametrize("cfg,steps", [("T0", 40), ("T1", 25)])
def test_algebra_modes_and_numpy_ekf_agree(cfg, steps):
    """as-written dense == structured == independent numpy EKF, every step, incl. association decisions."""
    st = make_stream(cfg, steps)
    a = Oracle(algebra=AS_WRITTEN, odom_model=st["model"])
    b = Oracle(algebra=STRUCTURED, odom_model=st["model"])
    c = NumpyEKF(odom_model=st["model"])
    for k in range(len(st["odom"])):
        o, cnt = st["odom"][k], int(st["obs_count"][k])
        xy = st["obs_xy"][k, :cnt]
        for f in (a, b):
            f.HandleOdometryMessage(*o)
            f.HandleObservationMessage(st["obs_time"][k], xy)
        c.handle_odometry(*o)
        c.handle_observation(st["obs_time"][k], xy)
        ma, mb = a.match_result(), b.match_result()
        for x, y, z in zip(ma, mb, c.last_match):
            assert np.array_equal(x, y) and np.array_equal(x.reshape(z.shape), z)
        assert np.abs(a.GetStateVector() - b.GetStateVector()).max() < 1e-12
        assert np.abs(a.GetStateVector() - c.mu).max() < 1e-11
        assert rel_fro(a.GetCoviarance(), b.GetCoviarance()) < 1e-12
        assert rel_fro(a.GetCoviarance(), c.sigma) < 1e-11
    N = st["N"]
    assert a.dim() == 3 + 2 * N          # every landmark created exactly once, no duplicates


def test_invariants_symmetry_psd_trace():
    st = make_stream("T1", 20)
    o = Oracle(algebra=STRUCTURED, odom_model=st["model"])
    for k in range(len(st["odom"])):
        od, cnt = st["odom"][k], int(st["obs_count"][k])
        o.HandleOdometryMessage(*od)
        # predict to the observation stamp first (m = 0 frame), then apply the update: trace must not grow
        o.HandleObservationMessage(st["obs_time"][k], np.zeros((0, 2), np.float32))
        tr0, n0 = np.trace(o.GetCoviarance()), o.dim()
        o.HandleObservationMessage(st["obs_time"][k], st["obs_xy"][k, :cnt])
        S = o.GetCoviarance()
        if o.dim() == n0:
            assert np.trace(S) <= tr0 * (1 + 1e-12)
        assert np.abs(S - S.T).max() <= 1e-12 * np.abs(S).max()
        assert np.linalg.eigvalsh((S + S.T) / 2).min() > -1e-10


def test_observation_permutation_only_permutes():
    """Shuffling a frame's observations changes nothing but the order of the new landmarks."""
    st = make_stream("T0", 6)
    a = Oracle(algebra=STRUCTURED)
    b = Oracle(algebra=STRUCTURED)
    rng = np.random.default_rng(5)
    for k in range(len(st["odom"])):
        od, cnt = st["odom"][k], int(st["obs_count"][k])
        xy = st["obs_xy"][k, :cnt]
        a.HandleOdometryMessage(*od); b.HandleOdometryMessage(*od)
        a.HandleObservationMessage(st["obs_time"][k], xy)
        b.HandleObservationMessage(st["obs_time"][k], xy[rng.permutation(cnt)])
    la = a.GetStateVector()[3:].reshape(-1, 2)
    lb = b.GetStateVector()[3:].reshape(-1, 2)
    assert la.shape == lb.shape
    order = [int(np.argmin(np.linalg.norm(lb - p, axis=1))) for p in la]
    assert sorted(order) == list(range(len(la)))
    assert np.abs(lb[order] - la).max() < 1e-6          # float32 rounding of new means, update order
    assert np.abs(a.GetStateVector()[:3] - b.GetStateVector()[:3]).max() < 1e-9


def test_negative_dt_and_stale_odometry():
    o = Oracle(algebra=AS_WRITTEN)
    n = NumpyEKF()
    o.HandleOdometryMessage(1.0, 0.5, 0, 0.1); n.handle_odometry(1.0, 0.5, 0, 0.1)
    o.HandleOdometryMessage(0.5, 9.0, 0, 9.0); n.handle_odometry(0.5, 9.0, 0, 9.0)   # stale: dropped (:211)
    assert o.GetLatestTime() == 1.0
    xy = np.array([[1.0, 1.0]], np.float32)
    o.HandleObservationMessage(0.9, xy); n.handle_observation(0.9, xy)                   # negative dt predict (:232)
    assert o.GetLatestTime() == 0.9
    assert np.abs(o.GetStateVector() - n.mu).max() < 1e-14
    assert rel_fro(o.GetCoviarance(), n.sigma) < 1e-13


def test_gps_pose_rows_match_numpy():
    st = make_stream("T0", 6)
    o = Oracle(algebra=AS_WRITTEN)
    n = NumpyEKF()
    for k in range(len(st["odom"])):
        od, cnt = st["odom"][k], int(st["obs_count"][k])
        xy = st["obs_xy"][k, :cnt]
        g = st["true_pose"][k] + np.array([0.01, -0.02, 0.003]) if k >= st["n_build"] else None
        o.HandleOdometryMessage(*od); n.handle_odometry(*od)
        o.HandleObservationMessage(st["obs_time"][k], xy, g); n.handle_observation(st["obs_time"][k], xy, g)
    assert np.abs(o.GetStateVector() - n.mu).max() < 1e-12
    assert rel_fro(o.GetCoviarance(), n.sigma) < 1e-11


def test_map_localisation_branch_and_txt_roundtrip(tmp_path):
    """Pre-loaded beacon map: sqrt(dᵀΣd) < 0.05 gate (:411,:420), A-only rows (:300), save → load."""
    st = make_stream("T0", 8)
    a = Oracle(algebra=AS_WRITTEN)
    for k in range(len(st["odom"])):
        od, cnt = st["odom"][k], int(st["obs_count"][k])
        a.HandleOdometryMessage(*od)
        a.HandleObservationMessage(st["obs_time"][k], st["obs_xy"][k, :cnt])
    base = str(tmp_path / "map")
    assert a.save_map_txt(base) == 0
    lines = open(base + ".txt").read().split("\n")
    assert lines[0].startswith(",") and lines[1].startswith(",")      # ros_node.cc:100,125 leading comma quirk
    # the reference's own loader chokes on the leading comma (std::stod("")): ours treats it as a no-op
    b = Oracle(algebra=AS_WRITTEN, map_path=base + ".txt")
    assert len(b.GetGlobalMap()[0]) == 0
    # a well-formed file (no leading comma)
    good = str(tmp_path / "good.txt")
    open(good, "w").write(lines[0][1:] + "\n" + lines[1][1:] + "\n")
    xy_saved = np.array([float(t) for t in lines[0][1:].split(",")]).reshape(-1, 2)
    for loader, algebra in ((0, AS_WRITTEN), (0, STRUCTURED), (1, AS_WRITTEN)):
        c = Oracle(algebra=algebra, map_path=good, map_loader=loader)
        mxy, mcov = c.GetGlobalMap()
        assert mxy.shape == (16, 2) and np.allclose(mxy, xy_saved.astype(np.float32))
        if loader == 0:
            assert np.all(mcov[:, 0, 0] > 0)
        else:   # REFERENCE loader reads the positions line (:90) and zeros past its end
            assert np.allclose(mcov[0].ravel(), xy_saved.ravel()[:4].astype(np.float64), rtol=1e-6)
            assert np.all(mcov[8:] == 0)
    # localisation run against the map with an inflated "covariance" so the (uninverted) gate can pass
    c = Oracle(algebra=AS_WRITTEN)
    d = Oracle(algebra=STRUCTURED)
    n = NumpyEKF()
    cov = np.tile(np.eye(2) * 0.05, (16, 1, 1))
    c.set_map(xy_saved, cov); d.set_map(xy_saved, cov)
    n.map_xy, n.map_cov = xy_saved.astype(np.float32), cov
    got_map_match = False
    for k in range(st["n_build"], len(st["odom"])):
        od, cnt = st["odom"][k], int(st["obs_count"][k])
        od = od.copy(); od[0] -= st["odom"][st["n_build"]][0] - 0.01
        t_obs = od[0] + 0.01
        xy = st["obs_xy"][k, :cnt]
        for f in (c, d):
            f.HandleOdometryMessage(*od); f.HandleObservationMessage(t_obs, xy)
        n.handle_odometry(*od); n.handle_observation(t_obs, xy)
        got_map_match |= len(c.match_result()[1]) > 0
        assert np.abs(c.GetStateVector() - n.mu).max() < 1e-11
        assert np.abs(c.GetStateVector() - d.GetStateVector()).max() < 1e-11
        assert rel_fro(c.GetCoviarance(), n.sigma) < 1e-10
    assert got_map_match


def test_predict_state_is_non_mutating():
    st = make_stream("T0", 3)
    o = Oracle(algebra=AS_WRITTEN)
    n = NumpyEKF()
    for k in range(len(st["odom"])):
        od, cnt = st["odom"][k], int(st["obs_count"][k])
        o.HandleOdometryMessage(*od); n.handle_odometry(*od)
        o.HandleObservationMessage(st["obs_time"][k], st["obs_xy"][k, :cnt]); n.handle_observation(st["obs_time"][k], st["obs_xy"][k, :cnt])
    before = o.GetStateVector().copy()
    mu, sig = o.PredictState(o.GetLatestTime() + 0.25)
    mu_n, sig_n = n.predict_state(n.time + 0.25)
    assert np.array_equal(before, o.GetStateVector())
    assert np.abs(mu - mu_n).max() < 1e-13 and rel_fro(sig, sig_n) < 1e-13


def test_oracle_under_asan_and_ubsan(tmp_path):
    """Host sanitizers on the C restatement: `make -C oracle asan` (-fsanitize=address,undefined), a T0 and a T1
    stream in both algebra modes plus the map text round trip, in a child process with the sanitizer runtime preloaded."""
    import shutil
    import subprocess
    import sys
    if not os.path.exists("/usr/bin/gcc"):
        pytest.skip("/usr/bin/gcc not available")
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan.so not found")
    subprocess.run(["make", "-s", "-C", os.path.join(here, "oracle"), "asan"], check=True)
    lib = os.path.join(here, "oracle", "_build", "liboracle_asan.so")
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan.so not found")
    code = f"""
import sys, ctypes
sys.path.insert(0, {here!r})
from oracle import pyoracle
from oracle.pyoracle import AS_WRITTEN, STRUCTURED, Oracle
from reflector_ekf_slam_b200.synth import make_stream
lib = ctypes.CDLL({lib!r})
pyoracle._bind(lib, ref=False)
for cfg, steps in (("T0", 12), ("T1", 6)):
    st = make_stream(cfg, steps)
    for alg in (AS_WRITTEN, STRUCTURED):
        o = Oracle(algebra=alg, lib=lib, odom_model=st["model"])
        for k in range(len(st["odom"])):
            o.HandleOdometryMessage(*st["odom"][k])
            o.HandleObservationMessage(st["obs_time"][k], st["obs_xy"][k, : st["obs_count"][k]])
        o.PredictState(o.GetLatestTime() + 0.01)
        assert o.save_map_txt({str(tmp_path / "m")!r}) == 0
        o.load_map_txt({str(tmp_path / "m.txt")!r})
        o.close()
print("asan run complete")
"""
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    assert "asan run complete" in res.stdout
    assert "AddressSanitizer" not in res.stderr and "runtime error" not in res.stderr, res.stderr[-3000:]
