"""GPU parity tests (run with -m gpu on a B200): the CUDA engine, driven through the C ABI by the
reference-shaped host class, against the CPU oracle on the same seeded inputs, every step.

Tolerances are BASELINE.json's: means within 1e-4 m, covariance within 1e-5 relative Frobenius, and
association decisions (ReflectorMatchResult) identical."""
import numpy as np
import pytest

from helpers import MU_TOL_M, SIGMA_REL_FRO, compare_matches, compare_state, drive_engine, drive_oracle, rel_fro

pytestmark = pytest.mark.gpu

COV_MODES = {"tcgen05": 0, "f64": 1, "i8": 2}


def _make(cfg_stream, cov, **kw):
    from oracle.pyoracle import STRUCTURED, AS_WRITTEN, Oracle
    from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM
    st = cfg_stream
    algebra = AS_WRITTEN if st["N"] <= 64 else STRUCTURED
    ekf = ReflectorEKFSLAM(odom_model=st["model"], max_landmarks=max(st["N"], 8), max_observations=max(st["m"], 8),
                           cov_update=COV_MODES[cov], **kw)
    orc = Oracle(algebra=algebra, odom_model=st["model"])
    return ekf, orc


@pytest.mark.parametrize("cov", ["f64", "tcgen05", "i8"])
@pytest.mark.parametrize("cfg,steps", [("T0", 40), ("T1", 30)])
def test_small_streams_every_step(engine_lib, cfg, steps, cov):
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream(cfg, steps)
    ekf, orc = _make(st, cov)
    worst = (0.0, 0.0)
    for k in range(len(st["odom"])):
        drive_engine(ekf, st, k)
        drive_oracle(orc, st, k)
        compare_matches(ekf, orc, f"{cfg} step {k}")
        d = compare_state(ekf, orc, tag=f"{cfg}/{cov} step {k}")
        worst = (max(worst[0], d[0]), max(worst[1], d[1]))
    assert ekf.error_flags() == 0
    assert abs(ekf.GetLatestTime() - orc.GetLatestTime()) == 0.0
    print(f"{cfg}/{cov}: worst |dmu| {worst[0]:.2e} m, worst rel-Fro {worst[1]:.2e}")


@pytest.mark.parametrize("cov", ["f64", "tcgen05", "i8"])
def test_c2_stream(engine_lib, cov):
    """Config C2 (N=256, m=50): map building + 40 steady steps, Σ compared every 5th step."""
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("C2", 40)
    ekf, orc = _make(st, cov)
    worst = (0.0, 0.0)
    for k in range(len(st["odom"])):
        drive_engine(ekf, st, k)
        drive_oracle(orc, st, k)
        compare_matches(ekf, orc, f"C2 step {k}")
        d = compare_state(ekf, orc, check_sigma=(k % 5 == 0 or k == len(st["odom"]) - 1), tag=f"C2/{cov} step {k}")
        worst = (max(worst[0], d[0]), max(worst[1], d[1]))
    assert ekf.dim() == 3 + 2 * 256 and ekf.error_flags() == 0
    print(f"C2/{cov}: worst |dmu| {worst[0]:.2e} m, worst rel-Fro {worst[1]:.2e}")


@pytest.mark.parametrize("cov", ["f64", "tcgen05", "i8"])
def test_c3_warm_start_and_steps(engine_lib, cov):
    """Config C3 (N=1024, m=100, the headline size): the oracle builds the map, the snapshot is injected
    with rekf_set_state, then 6 steady steps are compared."""
    from oracle.pyoracle import STRUCTURED, Oracle
    from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("C3", 6)
    orc = Oracle(algebra=STRUCTURED)
    for k in range(st["n_build"]):
        drive_oracle(orc, st, k)
    t, mu, sig = orc.GetState()
    assert mu.size == 3 + 2 * 1024
    ekf = ReflectorEKFSLAM(max_landmarks=1024, max_observations=100, cov_update=COV_MODES[cov])
    ekf.set_state(t, st["odom"][st["n_build"] - 1][1:4], mu, sig)
    assert rel_fro(ekf.GetCoviarance(), sig) < 1e-15 and np.array_equal(ekf.GetStateVector(), mu)
    for k in range(st["n_build"], len(st["odom"])):
        drive_engine(ekf, st, k)
        drive_oracle(orc, st, k)
        compare_matches(ekf, orc, f"C3 step {k}")
        d = compare_state(ekf, orc, tag=f"C3/{cov} step {k}")
    sp, _, nw = ekf.match_result()
    assert len(sp) == 100 and len(nw) == 0
    print(f"C3/{cov}: final |dmu| {d[0]:.2e} m, rel-Fro {d[1]:.2e}")


@pytest.mark.parametrize("cov", ["f64", "i8"])
def test_c4_omni_full_size(engine_lib, cov):
    """Config C4 (N=4096, m=200, OMNI odometry, n=8195, r=400 — past the shared-memory-resident Cholesky): the
    structured oracle builds the map in 21 frames, the snapshot is injected, 3 steady steps are compared
    (association lists, means, the whole 8195x8195 covariance), plus exact symmetry of the result."""
    from oracle.pyoracle import STRUCTURED, Oracle
    from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM
    from reflector_ekf_slam_b200.synth import OMNI, make_stream
    st = make_stream("C4", 3)
    assert st["model"] == OMNI
    orc = Oracle(algebra=STRUCTURED, odom_model=OMNI)
    for k in range(st["n_build"]):
        drive_oracle(orc, st, k)
    t, mu, sig = orc.GetState()
    assert mu.size == 3 + 2 * 4096
    ekf = ReflectorEKFSLAM(odom_model=OMNI, max_landmarks=4096, max_observations=200, cov_update=COV_MODES[cov])
    ekf.set_state(t, st["odom"][st["n_build"] - 1][1:4], mu, sig)
    del sig
    for k in range(st["n_build"], len(st["odom"])):
        drive_engine(ekf, st, k)
        drive_oracle(orc, st, k)
        compare_matches(ekf, orc, f"C4 step {k}")
        d = compare_state(ekf, orc, check_sigma=(k == len(st["odom"]) - 1), tag=f"C4/{cov} step {k}")
    sp, _, nw = ekf.match_result()
    assert len(sp) == 200 and len(nw) == 0 and ekf.error_flags() == 0
    S = ekf.GetCoviarance()
    assert np.array_equal(S, S.T)
    print(f"C4/{cov}: final |dmu| {d[0]:.2e} m, rel-Fro {d[1]:.2e}")


def test_c3_long_run_drift_i8(engine_lib):
    """Config C3, default int8-slice tensor-core covariance update, 80 steady steps: the error against the fp64 oracle
    must stay inside the north-star bar over the RUN (a per-step bias would integrate), the covariance must stay
    exactly symmetric, its trace must have shrunk, and the association lists must agree every step."""
    from oracle.pyoracle import STRUCTURED, Oracle
    from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("C3", 80)
    orc = Oracle(algebra=STRUCTURED)
    for k in range(st["n_build"]):
        drive_oracle(orc, st, k)
    t, mu, sig = orc.GetState()
    ekf = ReflectorEKFSLAM(max_landmarks=1024, max_observations=100, cov_update=COV_MODES["i8"])
    ekf.set_state(t, st["odom"][st["n_build"] - 1][1:4], mu, sig)
    worst = (0.0, 0.0)
    for k in range(st["n_build"], len(st["odom"])):
        drive_engine(ekf, st, k)
        drive_oracle(orc, st, k)
        compare_matches(ekf, orc, f"C3 step {k}")
        if (k - st["n_build"]) % 20 == 19:
            d = compare_state(ekf, orc, tag=f"C3/i8 long run step {k}")
            worst = (max(worst[0], d[0]), max(worst[1], d[1]))
    S = ekf.GetCoviarance()
    assert np.array_equal(S, S.T) and ekf.error_flags() == 0
    assert np.trace(S) < np.trace(sig)                    # 80 updates of 100 landmarks each remove far more than the odometry noise adds
    print(f"C3/i8 80 steps: worst |dmu| {worst[0]:.2e} m, worst rel-Fro {worst[1]:.2e}")


def test_covariance_stays_exactly_symmetric_and_psd(engine_lib):
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("T1", 20)
    for cov, floor in (("i8", -1e-9), ("f64", -1e-12), ("tcgen05", -1e-7)):
        ekf, _ = _make(st, cov)
        for k in range(len(st["odom"])):
            drive_engine(ekf, st, k)
        S = ekf.GetCoviarance()
        assert np.array_equal(S, S.T), cov            # bit-exact symmetry in every mode
        assert np.linalg.eigvalsh(S).min() > floor, cov


def test_negative_dt_stale_odometry_empty_frames(engine_lib):
    from oracle.pyoracle import AS_WRITTEN, Oracle
    from reflector_ekf_slam_b200.engine import Observation, OdometryData, ReflectorEKFSLAM
    ekf = ReflectorEKFSLAM(max_landmarks=8, max_observations=8)
    orc = Oracle(algebra=AS_WRITTEN)
    def both_odom(*a):
        ekf.HandleOdometryMessage(OdometryData(*a)); orc.HandleOdometryMessage(*a)
    def both_obs(t, xy):
        ekf.HandleObservationMessage(Observation(t, xy)); orc.HandleObservationMessage(t, np.asarray(xy, np.float32).reshape(-1, 2))
    both_odom(1.0, 0.5, 0.0, 0.1)
    both_odom(0.5, 9.0, 0.0, 9.0)                      # stale → dropped (:211)
    assert ekf.GetLatestTime() == 1.0
    both_obs(0.9, [[1.0, 1.0], [2.0, -1.0]])           # negative dt (:232), two new landmarks in one frame (:354 quirk)
    both_obs(1.1, np.zeros((0, 2)))                    # empty frame: predict only (:235)
    both_obs(1.2, [[1.02, 0.97]])
    compare_matches(ekf, orc)
    compare_state(ekf, orc)
    assert ekf.GetLatestTime() == orc.GetLatestTime() == 1.2
    S = ekf.GetCoviarance()
    So = orc.GetCoviarance()
    assert abs(S[3, 5] - So[3, 5]) < 1e-5 * abs(So[3, 5]) and S[3, 5] != 0.0    # same-frame +Qt cross block (:354)


def test_capacity_flags(engine_lib):
    from reflector_ekf_slam_b200.engine import Observation, RekfError, ReflectorEKFSLAM
    ekf = ReflectorEKFSLAM(max_landmarks=2, max_observations=4)
    ekf.HandleObservationMessage(Observation(0.1, [[1, 0], [3, 0], [5, 0]]))    # third reflector does not fit
    assert ekf.dim() == 3 + 2 * 2
    assert ekf.error_flags() & 1
    with pytest.raises(RekfError):
        ekf.sync()
    with pytest.raises(RekfError):
        ekf.HandleObservationMessage(Observation(0.2, np.zeros((5, 2))))         # > max_observations


def test_predict_state_and_accessors(engine_lib):
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("T0", 4)
    ekf, orc = _make(st, "f64")
    for k in range(len(st["odom"])):
        drive_engine(ekf, st, k); drive_oracle(orc, st, k)
    t = ekf.GetLatestTime() + 0.3
    _, mu, sig = ekf.PredictState(t)
    mu_o, sig_o = orc.PredictState(t)
    assert np.abs(mu - mu_o).max() < 1e-12 and rel_fro(sig, sig_o) < 1e-12
    compare_state(ekf, orc)                             # non-mutating
    pose, cov = ekf.pose()
    S = orc.GetCoviarance()
    assert np.abs(pose - orc.GetStateVector()[:3]).max() < MU_TOL_M and rel_fro(cov, S[:3, :3]) < SIGMA_REL_FRO
    xy, blocks = ekf.landmarks()
    assert xy.shape == (16, 2) and rel_fro(blocks[3], S[9:11, 9:11]) < SIGMA_REL_FRO
    # device-side marker extraction (ros_node.cc:736-789): the 95 % ellipse of every landmark's 2x2 block, checked
    # against numpy's symmetric eigen-decomposition of the oracle's covariance, and as a reconstruction of the block
    mk = ekf.markers()
    assert mk.shape == (16, 5) and np.array_equal(mk[:, :2], xy)
    for j in range(16):
        blk = S[3 + 2 * j:5 + 2 * j, 3 + 2 * j:5 + 2 * j]
        w = np.linalg.eigvalsh(blk)
        assert abs(mk[j, 3] - 2 * np.sqrt(w[1] * 5.991)) < 1e-6 * mk[j, 3] and abs(mk[j, 4] - 2 * np.sqrt(w[0] * 5.991)) < 1e-6 * mk[j, 3]
        R = np.array([[np.cos(mk[j, 2]), -np.sin(mk[j, 2])], [np.sin(mk[j, 2]), np.cos(mk[j, 2])]])
        lam = (mk[j, 3:5] / 2) ** 2 / 5.991
        assert rel_fro(R @ np.diag(lam) @ R.T, blk) < 1e-5


def test_gps_pose_rows(engine_lib):
    from oracle.pyoracle import AS_WRITTEN, Oracle
    from reflector_ekf_slam_b200.engine import Observation, OdometryData, ReflectorEKFSLAM
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("T0", 6)
    ekf = ReflectorEKFSLAM(max_landmarks=16, max_observations=8, cov_update=1)
    orc = Oracle(algebra=AS_WRITTEN)
    for k in range(len(st["odom"])):
        od, cnt = st["odom"][k], int(st["obs_count"][k])
        g = st["true_pose"][k] + np.array([0.01, -0.02, 0.003]) if k >= st["n_build"] else None
        ekf.HandleOdometryMessage(OdometryData(*od)); orc.HandleOdometryMessage(*od)
        ekf.HandleObservationMessage(Observation(st["obs_time"][k], st["obs_xy"][k, :cnt], g))
        orc.HandleObservationMessage(st["obs_time"][k], st["obs_xy"][k, :cnt], g)
        compare_state(ekf, orc, tag=f"gps step {k}")


def test_map_localisation_and_txt_roundtrip(engine_lib, tmp_path):
    from oracle.pyoracle import AS_WRITTEN, Oracle
    from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM
    from reflector_ekf_slam_b200.synth import make_stream
    st = make_stream("T0", 8)
    ekf, orc = _make(st, "f64")
    for k in range(len(st["odom"])):
        drive_engine(ekf, st, k); drive_oracle(orc, st, k)
    base_g, base_o = str(tmp_path / "g"), str(tmp_path / "o")
    ekf.save_map_txt(base_g); orc.save_map_txt(base_o)
    tg, to = open(base_g + ".txt").read(), open(base_o + ".txt").read()
    num = lambda s: np.array([float(x) for x in s.replace("\n", ",").split(",") if x])
    assert tg.count("\n") == to.count("\n") == 2 and tg.startswith(",")
    assert np.allclose(num(tg), num(to), rtol=2e-5, atol=1e-9)            # both are 6-significant-digit text
    good = str(tmp_path / "good.txt")
    l0, l1 = to.split("\n")[:2]
    open(good, "w").write(l0[1:] + "\n" + l1[1:] + "\n")
    loc = ReflectorEKFSLAM(max_landmarks=16, max_observations=8, map_path=good, cov_update=1)
    ref = Oracle(algebra=AS_WRITTEN, map_path=good)
    mxy, mcov = loc.GetGlobalMap()
    oxy, ocov = ref.GetGlobalMap()
    assert np.array_equal(mxy, oxy) and np.array_equal(mcov, ocov) and len(mxy) == 16
    cov = np.tile(np.eye(2) * 0.05, (16, 1, 1))
    loc.set_map(oxy, cov); ref.set_map(oxy, cov)
    got = False
    from reflector_ekf_slam_b200.engine import Observation, OdometryData
    t0 = st["odom"][st["n_build"]][0] - 0.01
    for k in range(st["n_build"], len(st["odom"])):
        od, cnt = st["odom"][k].copy(), int(st["obs_count"][k])
        od[0] -= t0
        xy = st["obs_xy"][k, :cnt]
        loc.HandleOdometryMessage(OdometryData(*od)); ref.HandleOdometryMessage(*od)
        loc.HandleObservationMessage(Observation(od[0] + 0.01, xy)); ref.HandleObservationMessage(od[0] + 0.01, xy)
        compare_matches(loc, ref, f"map step {k}")
        compare_state(loc, ref, tag=f"map step {k}")
        got |= len(loc.match_result()[1]) > 0
    assert got


def test_batched_sessions_match_single_sessions_bitwise(engine_lib):
    """S sessions in one handle == the same sessions run alone: bit-identical μ and Σ."""
    from reflector_ekf_slam_b200.engine import EKFBatch, ReflectorEKFSLAM
    from reflector_ekf_slam_b200.synth import make_stream
    S = 3
    sts = [make_stream("T1", 6, session=s) for s in range(S)]
    batch = EKFBatch(S, odom_model=sts[0]["model"], max_landmarks=60, max_observations=12)
    singles = [ReflectorEKFSLAM(odom_model=sts[0]["model"], max_landmarks=60, max_observations=12) for _ in range(S)]
    for k in range(len(sts[0]["odom"])):
        batch.handle_odometry(np.stack([st["odom"][k] for st in sts]))
        batch.handle_observation(np.array([st["obs_time"][k] for st in sts]), np.stack([st["obs_xy"][k] for st in sts]),
                                 np.array([st["obs_count"][k] for st in sts]))
        for s in range(S):
            drive_engine(singles[s], sts[s], k)
    for s in range(S):
        assert np.array_equal(batch.mu(s), singles[s].GetStateVector())
        assert np.array_equal(batch.sigma(s), singles[s].GetCoviarance())


def test_replay_device_matches_host_path(engine_lib):
    """rekf_replay_device (device-resident inputs, optional CUDA graph) == message-by-message host path."""
    import torch
    from reflector_ekf_slam_b200.engine import EKFBatch
    from reflector_ekf_slam_b200.synth import make_stream
    S, T = 2, 8
    sts = [make_stream("T1", T, session=s) for s in range(S)]
    nb, m = sts[0]["n_build"], sts[0]["m"]
    outs = []
    for graphs in (0, 1):
        b = EKFBatch(S, odom_model=sts[0]["model"], max_landmarks=60, max_observations=12, use_graphs=graphs)
        ref = EKFBatch(S, odom_model=sts[0]["model"], max_landmarks=60, max_observations=12)
        for k in range(nb):
            for e in (b, ref):
                e.handle_odometry(np.stack([st["odom"][k] for st in sts]))
                e.handle_observation(np.array([st["obs_time"][k] for st in sts]), np.stack([st["obs_xy"][k] for st in sts]),
                                     np.array([st["obs_count"][k] for st in sts]))
        for k in range(nb, nb + T):
            ref.handle_odometry(np.stack([st["odom"][k] for st in sts]))
            ref.handle_observation(np.array([st["obs_time"][k] for st in sts]), np.stack([st["obs_xy"][k] for st in sts]))
        dev = torch.device("cuda:0")
        d_odom = torch.tensor(np.stack([st["odom"][nb:] for st in sts]), device=dev)
        d_time = torch.tensor(np.stack([st["obs_time"][nb:] for st in sts]), device=dev)
        d_xy = torch.tensor(np.stack([st["obs_xy"][nb:] for st in sts]), device=dev)
        d_pose = torch.zeros(S, T, 3, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        b.sync()
        b.replay_device(d_odom.data_ptr(), d_time.data_ptr(), d_xy.data_ptr(), T, m, d_pose.data_ptr())
        b.sync()
        for s in range(S):
            assert np.array_equal(b.mu(s), ref.mu(s))
            assert np.array_equal(b.sigma(s), ref.sigma(s))
            assert np.array_equal(d_pose[s, -1].cpu().numpy(), ref.mu(s)[:3])
        outs.append(d_pose.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("cov", ["i8", "f64"])
def test_shadow_trsm_bitwise_and_parity(engine_lib, cov, monkeypatch):
    """The TRSM that runs BESIDE the Cholesky (solve_ll.cuh: flags in global memory, programmatic launch chain) against the
    same kernel run in order (REKF_SHADOW=0: every flag is up before it starts): bit-identical state after every step, and
    both inside the north-star bar against the oracle.  N=120, m=40: r = 80 = three 32-row blocks (diagonal products, trailing
    updates, thread-private update storage, the gather one block ahead); then 3 steps at C3 (seven blocks, 68 tiles)."""
    from oracle.pyoracle import STRUCTURED, Oracle
    from reflector_ekf_slam_b200.engine import ReflectorEKFSLAM
    from reflector_ekf_slam_b200.synth import DIFF, make_stream

    def engines(**kw):
        out = []
        for shadow in ("0", "1"):
            monkeypatch.setenv("REKF_SHADOW", shadow)    # read when the handle is created
            out.append(ReflectorEKFSLAM(cov_update=COV_MODES[cov], **kw))
        monkeypatch.delenv("REKF_SHADOW")
        return out

    st = make_stream(None, 12, N=120, m=40, model=DIFF, seed=77)
    in_order, shadow = engines(odom_model=st["model"], max_landmarks=120, max_observations=40)
    orc = Oracle(algebra=STRUCTURED, odom_model=st["model"])
    for k in range(len(st["odom"])):
        drive_engine(in_order, st, k)
        drive_engine(shadow, st, k)
        drive_oracle(orc, st, k)
        assert np.array_equal(shadow.GetStateVector(), in_order.GetStateVector()), f"step {k}"
        if k % 3 == 0 or k == len(st["odom"]) - 1:
            assert np.array_equal(shadow.GetCoviarance(), in_order.GetCoviarance()), f"step {k}"
            compare_matches(shadow, orc, f"shadow step {k}")
            compare_state(shadow, orc, tag=f"shadow/{cov} step {k}")
    assert shadow.error_flags() == 0 and in_order.error_flags() == 0

    st = make_stream("C3", 3)
    in_order, shadow = engines(max_landmarks=1024, max_observations=100)
    for k in range(len(st["odom"])):
        drive_engine(in_order, st, k)
        drive_engine(shadow, st, k)
    assert np.array_equal(shadow.GetStateVector(), in_order.GetStateVector())
    assert np.array_equal(shadow.GetCoviarance(), in_order.GetCoviarance())
    assert shadow.error_flags() == 0 and len(shadow.match_result()[0]) == 100


@pytest.mark.parametrize("groups", [2, 3])
def test_pipeline_groups_bitwise(engine_lib, groups):
    """pipeline_groups > 1 (each group of sessions on its own stream, persistent SYRK fed from the atomic tile queue)
    produces bit-identical states to the single-group batch, through the host path, the device replay (with and
    without graphs) and the streaming pose read."""
    import torch
    from reflector_ekf_slam_b200.engine import EKFBatch
    from reflector_ekf_slam_b200.synth import make_stream
    S, T = 5, 8
    sts = [make_stream("T1", 2 * T, session=s) for s in range(S)]
    nb, m = sts[0]["n_build"], sts[0]["m"]
    kw = dict(odom_model=sts[0]["model"], max_landmarks=60, max_observations=12)
    ref = EKFBatch(S, **kw)
    engines = [ref] + [EKFBatch(S, pipeline_groups=groups, use_graphs=g, **kw) for g in (0, 1)]
    tickets = []
    for k in range(nb + T):
        for e in engines:
            e.handle_odometry(np.stack([st["odom"][k] for st in sts]))
            e.handle_observation(np.array([st["obs_time"][k] for st in sts]), np.stack([st["obs_xy"][k] for st in sts]),
                                 np.array([st["obs_count"][k] for st in sts]))
        tickets.append((engines[1].request_poses(), ref.poses().copy()))
        if len(tickets) > 4:
            t, want = tickets.pop(0)
            assert np.array_equal(engines[1].fetch_poses(t), want)
    for t, want in tickets:
        assert np.array_equal(engines[1].fetch_poses(t), want)
    dev = torch.device("cuda:0")
    d_odom = torch.tensor(np.stack([st["odom"][nb + T:] for st in sts]), device=dev)
    d_time = torch.tensor(np.stack([st["obs_time"][nb + T:] for st in sts]), device=dev)
    d_xy = torch.tensor(np.stack([st["obs_xy"][nb + T:] for st in sts]), device=dev)
    poses = []
    for e in engines:
        d_pose = torch.zeros(S, T, 3, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        e.replay_device(d_odom.data_ptr(), d_time.data_ptr(), d_xy.data_ptr(), T, m, d_pose.data_ptr())
        e.sync()
        poses.append(d_pose.cpu().numpy())
    # the one-call step (both messages, one copy, one graph per group) == the two message calls
    stepper = EKFBatch(S, pipeline_groups=groups, use_graphs=1, **kw)
    for k in range(nb + 2 * T):
        stepper.handle_step(np.stack([st["odom"][k] for st in sts]), np.array([st["obs_time"][k] for st in sts]),
                            np.stack([st["obs_xy"][k] for st in sts]), np.array([st["obs_count"][k] for st in sts]))
    for s in range(S):
        assert np.array_equal(stepper.mu(s), ref.mu(s)) and np.array_equal(stepper.sigma(s), ref.sigma(s))
    for e, p in zip(engines[1:], poses[1:]):
        assert np.array_equal(p, poses[0])
        for s in range(S):
            assert np.array_equal(e.mu(s), ref.mu(s)) and np.array_equal(e.sigma(s), ref.sigma(s))
            for a, b in zip(e.match_result(s), ref.match_result(s)):
                assert np.array_equal(a, b)
