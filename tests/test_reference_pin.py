"""The pin: the reference's OWN translation units (reflector_ekf_slam.cc / reflector_ekf_slam_gps.cc, compiled
unmodified into oracle/_ref/ against oracle/shim — see oracle/ref_abi.cc, `make -C oracle ref`) against the C
restatement (oracle/rekf_oracle.c), every step, on every BASELINE configuration that the as-written algebra can
afford on the CPU, plus the shipped rosbag.  No GPU.

Tolerance 1e-12 (the two differ only in floating-point summation order inside the dense products)."""
import os

import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import AS_WRITTEN, STRUCTURED, Oracle
from reflector_ekf_slam_b200.synth import make_stream

from helpers import drive_oracle, rel_fro

pytestmark = pytest.mark.skipif(not pyoracle.ref_available(), reason="oracle/_ref not built and /root/reference absent")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12


def Reference(**kw):
    return pyoracle.Reference(**kw)


def assert_same(a, r, tag, tol=TOL, sigma=True):
    assert a.dim() == r.dim(), tag
    assert a.GetLatestTime() == r.GetLatestTime(), tag
    for x, y, name in zip(a.match_result(), r.match_result(), ("state", "map", "new")):
        assert np.array_equal(x, y), f"{tag}: {name} match lists differ"
    dmu = float(np.abs(a.GetStateVector() - r.GetStateVector()).max())
    assert dmu < tol, f"{tag}: dmu {dmu:.3e}"
    if sigma:
        ds = rel_fro(a.GetCoviarance(), r.GetCoviarance())
        assert ds < tol, f"{tag}: Sigma rel Fro {ds:.3e}"


def test_shim_products_inverse_and_laziness():
    """The Eigen stand-in itself: blocked GEMM incl. transposed operands, partial-pivot LU inverse, lazy
    re-evaluation — through a tiny C++ probe compiled against oracle/shim."""
    import subprocess
    import tempfile
    src = r'''
#include <Eigen/Dense>
#include <cstdio>
#include <cmath>
int main() {
  const int n = 157, r = 44;
  Eigen::MatrixXd A = Eigen::MatrixXd::Zero(n, n), H = Eigen::MatrixXd::Zero(r, n), Q = Eigen::MatrixXd::Zero(r, r);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1u << 24) - 0.5; };
  for (int j = 0; j < n; ++j) for (int i = 0; i <= j; ++i) { double v = rnd(); A(i, j) = v; A(j, i) = v; }
  for (int i = 0; i < n; ++i) A(i, i) += n;
  for (int j = 0; j < n; ++j) for (int i = 0; i < r; ++i) H(i, j) = rnd();
  for (int i = 0; i < r; ++i) Q(i, i) = 0.01;
  // naive references
  double worst = 0;
  Eigen::MatrixXd P = H * A * H.transpose() + Q;
  for (int i = 0; i < r; ++i) for (int j = 0; j < r; ++j) {
    double acc = (i == j) ? 0.01 : 0.0;
    for (int k = 0; k < n; ++k) { double t = 0; for (int l = 0; l < n; ++l) t += H(i, l) * A(l, k); acc += t * H(j, k); }
    worst = std::fmax(worst, std::fabs(acc - P(i, j)));
  }
  Eigen::MatrixXd I = P * P.inverse();
  double worst_inv = 0;
  for (int i = 0; i < r; ++i) for (int j = 0; j < r; ++j) worst_inv = std::fmax(worst_inv, std::fabs(I(i, j) - (i == j)));
  // laziness: the expression sees later changes of its operands (it is evaluated at use, :305-308)
  const auto lazy = A * H.transpose();
  Eigen::MatrixXd before = lazy;
  A(0, 0) += 1.0;
  Eigen::MatrixXd after = lazy;
  const double moved = std::fabs(after(0, 0) - before(0, 0) - H(0, 0));
  // 1x1 fixed-size product converts to a scalar (:411)
  Eigen::Vector2f d(3.f, 4.f);
  const auto dd = d.cast<double>().transpose();
  const double dist = std::sqrt(dd * dd.transpose());
  std::printf("%.3e %.3e %.3e %.17g\n", worst, worst_inv, moved, dist);
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as td:
        cc = os.path.join(td, "probe.cc")
        open(cc, "w").write(src)
        exe = os.path.join(td, "probe")
        subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "oracle", "shim"), cc, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    worst, worst_inv, moved, dist = map(float, out)
    assert worst < 1e-9 and worst_inv < 1e-10 and moved < 1e-12 and dist == 5.0


@pytest.mark.parametrize("cfg,steps", [("T0", 40), ("T1", 30)])
def test_ref_equals_restatement_small_every_step(cfg, steps):
    st = make_stream(cfg, steps)
    a = Oracle(algebra=AS_WRITTEN, odom_model=st["model"])
    r = Reference(odom_model=st["model"])
    for k in range(len(st["odom"])):
        drive_oracle(a, st, k)
        drive_oracle(r, st, k)
        assert_same(a, r, f"{cfg} step {k}")
    assert r.dim() == 3 + 2 * st["N"]


def test_ref_equals_restatement_c2_full_stream_and_structured():
    """C2 (N=256, m=50): map building through the reference's own augmentation + 6 steady-state steps;
    reference == as-written restatement == structured restatement at every step."""
    st = make_stream("C2", 6)
    a = Oracle(algebra=AS_WRITTEN, odom_model=st["model"])
    s = Oracle(algebra=STRUCTURED, odom_model=st["model"])
    r = Reference(odom_model=st["model"], fast=True)
    for k in range(len(st["odom"])):
        for f in (a, s, r):
            drive_oracle(f, st, k)
        assert_same(a, r, f"C2 step {k} as-written vs reference")
        assert_same(s, r, f"C2 step {k} structured vs reference")
    assert r.dim() == 515


def test_ref_equals_restatement_c3_steps():
    """C3 (N=1024, m=100, n=2051): 3 steady-state steps from the map-building snapshot; reference's own code
    (~2-6 s per step) == as-written restatement == structured restatement."""
    st = make_stream("C3", 3)
    s = Oracle(algebra=STRUCTURED, odom_model=st["model"])
    nb = st["n_build"]
    for k in range(nb):
        drive_oracle(s, st, k)
    t, mu, S = s.GetState()
    vt = st["odom"][nb - 1][1:4]
    a = Oracle(algebra=AS_WRITTEN, odom_model=st["model"], native=True)
    r = Reference(odom_model=st["model"], fast=True)
    for f in (a, r):
        f.set_state(t, vt, mu, S)
    for k in range(nb, nb + 3):
        for f in (a, s, r):
            drive_oracle(f, st, k)
        assert_same(a, r, f"C3 step {k} as-written vs reference")
        assert_same(s, r, f"C3 step {k} structured vs reference")
    assert r.dim() == 2051 and len(r.match_result()[0]) == 100


def test_ref_gps_class_equals_restatement():
    """ekf::ReflectorEKFSLAMGPS (reflector_ekf_slam_gps.cc:305-340): pose pseudo-measurement rows."""
    st = make_stream("T0", 30)
    a = Oracle(algebra=AS_WRITTEN, odom_model=st["model"])
    r = Reference(odom_model=st["model"], gps=True)
    rng = np.random.default_rng(7)
    for k in range(len(st["odom"])):
        o, c = st["odom"][k], int(st["obs_count"][k])
        gps = None
        if k % 3 == 1:
            gps = st["true_pose"][k] + rng.normal(0, [0.03, 0.03, 0.01])
            if k % 2:
                gps[2] += 2 * np.pi       # the residual is wrapped through a quaternion (:326-328)
        for f in (a, r):
            f.HandleOdometryMessage(*o)
            f.HandleObservationMessage(st["obs_time"][k], st["obs_xy"][k, :c], gps_pose=gps)
        assert_same(a, r, f"gps step {k}")


def test_ref_predict_state_stale_odometry_negative_dt_empty_frames():
    st = make_stream("T1", 12)
    a = Oracle(algebra=AS_WRITTEN, odom_model=st["model"])
    r = Reference(odom_model=st["model"])
    for k in range(len(st["odom"])):
        drive_oracle(a, st, k)
        drive_oracle(r, st, k)
    t = a.GetLatestTime()
    for f in (a, r):
        f.HandleOdometryMessage(t - 0.5, 9.0, 9.0, 9.0)                       # stale: dropped (:211)
        f.HandleObservationMessage(t - 0.01, np.zeros((0, 2), np.float32))    # negative dt, empty cloud (:232-236)
        f.HandleObservationMessage(t - 0.02, st["obs_xy"][-1, :5])            # negative dt with an update
    assert_same(a, r, "edge cases")
    for tq in (t + 0.3, t - 0.1):
        mu_a, S_a = a.PredictState(tq)
        mu_r, S_r = r.PredictState(tq)
        assert np.abs(mu_a - mu_r).max() < TOL and rel_fro(S_a, S_r) < TOL


def test_ref_map_loader_and_map_localisation(tmp_path):
    """LoadMapFromTxtFile (:43-95) — the reference's loader reads the covariances from line 1 (:87-91) — and the
    beacon-map branch of ReflectorMatch / the update (:401-425, :279-303)."""
    from reflector_ekf_slam_b200._abi import REKF_MAP_LOADER_REFERENCE
    st = make_stream("T0", 30)
    b = Oracle(algebra=STRUCTURED, odom_model=st["model"])
    for k in range(st["n_build"] + 4):
        drive_oracle(b, st, k)
    # a well-formed two-line file (what SaveReflectorResult writes when the beacon map is non-empty and the state
    # has no landmarks, ros_node.cc:86-97,112-123): no leading comma
    mu, S = b.GetStateVector(), b.GetCoviarance()
    lm = mu[3:].reshape(-1, 2)
    path = str(tmp_path / "map.txt")
    with open(path, "w") as f:
        f.write(",".join("%g,%g" % (x, y) for x, y in lm) + "\n")
        f.write(",".join("%g,%g,%g,%g" % (S[3 + 2 * i, 3 + 2 * i], S[3 + 2 * i, 4 + 2 * i], S[4 + 2 * i, 3 + 2 * i],
                                          S[4 + 2 * i, 4 + 2 * i]) for i in range(len(lm))) + "\n")
    a = Oracle(algebra=AS_WRITTEN, odom_model=st["model"], map_path=path, map_loader=REKF_MAP_LOADER_REFERENCE)
    r = Reference(odom_model=st["model"], map_path=path)
    xy_a, cov_a = a.GetGlobalMap()
    xy_r, cov_r = r.GetGlobalMap()
    assert xy_a.shape == xy_r.shape == (16, 2)
    assert np.array_equal(xy_a, xy_r)
    # :87-91 indexes result[0] (32 numbers) with 4*i+k for i < 16: the first 8 "covariances" are landmark
    # coordinates, the rest is an out-of-bounds heap read in the reference (undefined); the restatement reads 0.0
    assert np.array_equal(cov_a[:8], cov_r[:8])
    assert np.array_equal(cov_a[:8].reshape(-1), lm.astype(np.float32).astype(np.float64).reshape(-1)[:32]) or \
        np.allclose(cov_a[:8].reshape(-1), [float("%g" % v) for v in lm.reshape(-1)[:32]])
    # with the covariances the author meant (tiny variances, sqrt(d' S d) < 0.05 is easy to meet) beacons do match
    cov = np.tile(np.eye(2) * 1e-3, (16, 1, 1))
    a2 = Oracle(algebra=AS_WRITTEN, odom_model=st["model"])
    r2 = Reference(odom_model=st["model"])
    for f in (a2, r2):
        f.set_map(xy_r, cov)
    n_map = 0
    for k in range(len(st["odom"])):
        drive_oracle(a2, st, k)
        drive_oracle(r2, st, k)
        assert_same(a2, r2, f"beacon step {k}")
        n_map += len(r2.match_result()[1])
    assert n_map > 20


def test_ref_bag_replay_equals_restatement_and_finds_seven_reflectors():
    """Config C1: the reference's own EKF on the observation stream of its shipped rosbag."""
    from test_bag_replay import EXPECTED_LANDMARKS, load_bag, replay
    bag = load_bag()
    t0 = float(bag["init_time"])
    a = Oracle(algebra=AS_WRITTEN, init_time=t0)
    r = Reference(init_time=t0)
    replay(a.HandleOdometryMessage, a.HandleObservationMessage, bag)
    replay(r.HandleOdometryMessage, r.HandleObservationMessage, bag)
    assert_same(a, r, "bag", tol=1e-10)
    lm = r.GetStateVector()[3:].reshape(-1, 2)
    assert lm.shape == (7, 2)
    for e in EXPECTED_LANDMARKS:
        assert np.linalg.norm(lm - e, axis=1).min() < 0.05


def test_ref_loader_aborts_on_the_file_its_own_node_saves(tmp_path):
    """Node::SaveReflectorResult writes a leading comma when the beacon map is empty (ros_node.cc:101), and the
    reference's LoadMapFromTxtFile then calls std::stod("") (:58), which throws and terminates the process.
    This documents the reference's behaviour; the restatement and the engine treat such a file as a no-op
    (defined deviation)."""
    import subprocess
    import sys
    st = make_stream("T0", 2)
    b = Oracle(algebra=STRUCTURED, odom_model=st["model"])
    for k in range(st["n_build"]):
        drive_oracle(b, st, k)
    base = str(tmp_path / "saved")
    assert b.save_map_txt(base) == 0
    assert open(base + ".txt").read().startswith(",")
    code = ("import sys; sys.path.insert(0, %r); from oracle import pyoracle; "
            "pyoracle.Reference(map_path=%r); print('survived')" % (ROOT, base + ".txt"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert res.returncode != 0 and "survived" not in res.stdout and "stod" in res.stderr
    a = Oracle(algebra=AS_WRITTEN, map_path=base + ".txt")
    assert a.GetGlobalMap()[0].shape == (0, 2)
