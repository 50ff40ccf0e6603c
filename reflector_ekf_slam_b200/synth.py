"""Synthetic 2-D reflector + odometry streams for the BASELINE.json configurations.

The reference ships one rosbag and no synthetic workload; the streams here are the
ones SURVEY.md §8(d) specifies, with two details fixed so that the steady state really is
"N landmarks in the state, m observed per step, all matched" (the configuration the metric is
quoted on):

* the landmark lattice is centred on the start pose (the robot starts in the middle of the map);
* during the map-building warm start the *odometry measurements* are noise-free (the filter still
  applies its odometry noise model, so Σ is a genuine EKF covariance with all cross-correlations,
  including the reference's same-frame `+Qt` quirk, reflector_ekf_slam.cc:354).  With noisy odometry
  and no corrections during those frames, heading drift × 45 m range would initialise far landmarks
  > 0.6 m (the gate at :446) from where they are later re-observed, and the state would fill with
  duplicates instead of reaching the configured N.

A stream for one session is a dict of numpy arrays:
    odom      (T, 4)  float64   time, vx, vy, wz      — one HandleOdometryMessage per step
    obs_time  (T,)    float64                         — one HandleObservationMessage per step,
    obs_xy    (T, m, 2) float32                         stamped dt/2 after the odometry message
    obs_count (T,)    int32     valid observations in the frame (< m only in the last build frame)
    n_build   int               the first n_build steps are the map-building phase
    landmarks (N, 2)  float64   ground truth (not visible to the filter)
    true_pose (T, 3)  float64   ground-truth pose at each observation stamp
"""
import math

import numpy as np

DIFF, OMNI = 0, 1
BASE_SEED = 20260317

# name -> (config index for the seed, N landmarks, m observed per step, odometry model)
CONFIGS = {
    "T0": (0, 16, 4, DIFF),      # tiny: as-written oracle in milliseconds (tests)
    "T1": (10, 60, 12, OMNI),    # small omni case (tests)
    "C2": (2, 256, 50, DIFF),
    "C3": (3, 1024, 100, DIFF),  # the headline configuration
    "C4": (4, 4096, 200, OMNI),
}

SIGMA_V, SIGMA_W, SIGMA_Z = 0.05, 0.08, 0.05   # launch/slam.launch:21-23
ODOM_HZ = 30.0


def _commanded_velocity(model, t):
    w = 0.05 * math.sin(0.1 * t)
    if model == DIFF:
        return 0.5, 0.0, w
    return 0.5, 0.2 * math.cos(0.05 * t), w


def _integrate(pose, vx, vy, w, dt, substeps=8):
    x, y, th = pose
    h = dt / substeps
    for _ in range(substeps):
        thm = th + 0.5 * w * h
        x += (vx * math.cos(thm) - vy * math.sin(thm)) * h
        y += (vx * math.sin(thm) + vy * math.cos(thm)) * h
        th += w * h
    return x, y, th


def lattice(N, rng):
    """N landmarks on a jittered square lattice, pitch 2.0 m, jitter U(-0.3, 0.3) per axis → pairwise
    separation >= 1.4 m > 2 x the 0.6 m association gate.  Centred on the origin."""
    side = int(math.ceil(math.sqrt(N)))
    idx = np.arange(side * side)[:N]
    gx = (idx % side).astype(np.float64)
    gy = (idx // side).astype(np.float64)
    pts = np.stack([gx, gy], 1) * 2.0
    pts -= (side - 1) * 1.0
    pts += rng.uniform(-0.3, 0.3, size=pts.shape)
    return pts


def make_stream(config="C3", steps=64, session=0, N=None, m=None, model=None, seed=None):
    """Build one session's stream: the map-building phase (ceil(N/m) frames) followed by `steps`
    steady-state steps."""
    if config is not None:
        ci, cN, cm, cmodel = CONFIGS[config]
    else:
        ci, cN, cm, cmodel = 99, N, m, model
    N = cN if N is None else N
    m = cm if m is None else m
    model = cmodel if model is None else model
    if seed is None:
        seed = BASE_SEED + ci + 1000 * session
    rng = np.random.default_rng(seed)
    lms = lattice(N, rng)
    n_build = int(math.ceil(N / m))
    T = n_build + steps
    dt = 1.0 / ODOM_HZ
    odom = np.zeros((T, 4))
    obs_time = np.zeros(T)
    obs_xy = np.zeros((T, m, 2), np.float32)
    obs_count = np.zeros(T, np.int32)
    true_pose = np.zeros((T, 3))
    pose = (0.0, 0.0, 0.0)
    t_obs_prev = 0.0
    for k in range(T):
        t_odom = (k + 1) * dt
        t_obs = t_odom + dt / 2
        vx, vy, w = _commanded_velocity(model, t_odom)
        pose = _integrate(pose, vx, vy, w, t_obs - t_obs_prev)   # velocity k holds over (t_obs[k-1], t_obs[k]]
        t_obs_prev = t_obs
        building = k < n_build
        nv, nw = (0.0, 0.0) if building else (rng.normal(0, SIGMA_V), rng.normal(0, SIGMA_W))
        nvy = 0.0 if (building or model == DIFF) else rng.normal(0, SIGMA_V)
        odom[k] = (t_odom, vx + nv, vy + nvy, w + nw)
        obs_time[k] = t_obs
        true_pose[k] = pose
        if building:
            ids = np.arange(k * m, min((k + 1) * m, N))
        else:
            d2 = np.sum((lms - np.array(pose[:2])) ** 2, axis=1)
            ids = np.argsort(d2, kind="stable")[:m]
        ids = rng.permutation(ids)
        c, s = math.cos(pose[2]), math.sin(pose[2])
        d = lms[ids] - np.array(pose[:2])
        local = np.stack([d[:, 0] * c + d[:, 1] * s, -d[:, 0] * s + d[:, 1] * c], 1)
        local += rng.normal(0, SIGMA_Z, size=local.shape)
        obs_xy[k, : len(ids)] = local.astype(np.float32)
        obs_count[k] = len(ids)
    return {
        "config": config, "N": N, "m": m, "model": model, "seed": seed, "n_build": n_build,
        "odom": odom, "obs_time": obs_time, "obs_xy": obs_xy, "obs_count": obs_count,
        "landmarks": lms, "true_pose": true_pose, "dt": dt,
    }


def stream_checksum(stream):
    """Order-sensitive checksum of a stream's inputs (used to verify the per-rank scatter)."""
    import zlib
    h = 0
    for key in ("odom", "obs_time", "obs_xy", "obs_count"):
        h = zlib.crc32(np.ascontiguousarray(stream[key]).tobytes(), h)
    return h
