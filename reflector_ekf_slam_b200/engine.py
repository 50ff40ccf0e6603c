"""Python host side over the C ABI (include/rekf.h) — mirrors the reference's operator interface.

`ReflectorEKFSLAM` keeps the method names, argument meaning and (silent) error behaviour of
ekf::ReflectorEKFSLAMInterface (reference include/reflector_ekf_slam/ekf_slam_interface.h:50-67) so the
parity tests read like tests of the reference class.  `EKFBatch` is the batched-session form
(S independent filters advancing through the same launches).  Everything goes through ctypes into
librekf_b200.so; there is no Python / CPU implementation behind these classes — if the library is
missing or no sm_100 GPU is visible, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import RekfOptions, make_options  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librekf_b200.so")
_lib = None

# every symbol include/rekf.h declares (tests check that the built library exports all of them)
ABI_SYMBOLS = [
    "rekf_default_options", "rekf_create", "rekf_create_batch", "rekf_destroy", "rekf_last_error", "rekf_version",
    "rekf_sessions", "rekf_handle_odometry", "rekf_handle_observation", "rekf_handle_imu",
    "rekf_batch_handle_odometry", "rekf_batch_handle_observation", "rekf_replay_device", "rekf_dim", "rekf_time",
    "rekf_get_mu", "rekf_get_pose", "rekf_batch_get_pose", "rekf_get_landmarks", "rekf_get_sigma", "rekf_get_match_result",
    "rekf_predict_state", "rekf_set_state", "rekf_set_map", "rekf_get_map", "rekf_load_map_txt", "rekf_save_map_txt",
    "rekf_sync", "rekf_stream", "rekf_timer_start", "rekf_timer_stop", "rekf_profile_enable", "rekf_profile_read",
    "rekf_launch_count", "rekf_device_error_flags", "rekf_debug_copy", "rekf_batch_request_poses", "rekf_batch_fetch_poses",
    "rekf_get_markers", "rekf_batch_handle_step", "rekf_get_state", "rekf_host_register", "rekf_host_unregister", "rekf_get_counters",
]


class RekfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rekf error {code}: {msg}")
        self.code = code


def load_library(path=None):
    """dlopen librekf_b200.so (built in-tree by reflector_ekf_slam_b200/build.py).  Fails loudly."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or _LIB_PATH
    if not os.path.exists(p):
        raise RekfError(_abi.REKF_ERR_UNSUPPORTED, f"{p} not built — run `python -m reflector_ekf_slam_b200.build` (no CPU fallback exists)")
    lib = C.CDLL(p)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    P = C.POINTER
    sig = {
        "rekf_default_options": (None, [P(RekfOptions)]),
        "rekf_create": (i, [P(RekfOptions), P(vp)]),
        "rekf_create_batch": (i, [P(RekfOptions), i, P(vp)]),
        "rekf_destroy": (i, [vp]),
        "rekf_last_error": (C.c_char_p, [vp]),
        "rekf_version": (C.c_char_p, []),
        "rekf_sessions": (i, [vp]),
        "rekf_handle_odometry": (i, [vp, d, d, d, d]),
        "rekf_handle_observation": (i, [vp, d, vp, i, vp]),
        "rekf_handle_imu": (i, [vp, d]),
        "rekf_batch_handle_odometry": (i, [vp, vp]),
        "rekf_batch_handle_observation": (i, [vp, vp, vp, vp, i]),
        "rekf_replay_device": (i, [vp, vp, vp, vp, i, i, vp]),
        "rekf_dim": (i, [vp, i]),
        "rekf_time": (i, [vp, i, P(d)]),
        "rekf_get_mu": (i, [vp, i, vp, i, P(i)]),
        "rekf_get_pose": (i, [vp, i, vp, vp]),
        "rekf_batch_get_pose": (i, [vp, vp]),
        "rekf_get_landmarks": (i, [vp, i, vp, vp, i, P(i)]),
        "rekf_get_sigma": (i, [vp, i, vp, i]),
        "rekf_get_match_result": (i, [vp, i, vp, P(i), vp, P(i), vp, P(i), i]),
        "rekf_predict_state": (i, [vp, i, d, vp, i, vp, i]),
        "rekf_set_state": (i, [vp, i, d, vp, vp, i, vp, i]),
        "rekf_set_map": (i, [vp, vp, vp, i]),
        "rekf_get_map": (i, [vp, vp, vp, i, P(i)]),
        "rekf_load_map_txt": (i, [vp, C.c_char_p]),
        "rekf_save_map_txt": (i, [vp, i, C.c_char_p]),
        "rekf_sync": (i, [vp]),
        "rekf_stream": (vp, [vp]),
        "rekf_timer_start": (i, [vp]),
        "rekf_timer_stop": (i, [vp, P(C.c_float)]),
        "rekf_profile_enable": (i, [vp, i]),
        "rekf_profile_read": (i, [vp, P(C.c_char_p), P(d), P(i), i, P(i)]),
        "rekf_launch_count": (C.c_int64, [vp]),
        "rekf_device_error_flags": (i, [vp, i, P(i)]),
        "rekf_debug_copy": (i, [vp, i, C.c_char_p, vp, C.c_size_t]),
        "rekf_batch_request_poses": (i, [vp, P(C.c_int64)]),
        "rekf_batch_fetch_poses": (i, [vp, C.c_int64, vp]),
        "rekf_get_markers": (i, [vp, i, vp, i, P(i)]),
        "rekf_batch_handle_step": (i, [vp, vp, vp, vp, vp, i]),
        "rekf_get_state": (i, [vp, i, i, P(d), vp, vp, i, P(i), P(i)]),
        "rekf_host_register": (i, [vp, vp, C.c_size_t]),
        "rekf_host_unregister": (i, [vp, vp]),
        "rekf_get_counters": (i, [vp, i, P(C.c_int64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EKFBatch:
    """S independent filters on one GPU behind one handle (rekf_create_batch)."""

    def __init__(self, sessions=1, options=None, **kw):
        self.lib = load_library()
        self.options = options if options is not None else make_options(**kw)
        self.S = int(sessions)
        self.h = C.c_void_p()
        rc = self.lib.rekf_create_batch(C.byref(self.options), self.S, C.byref(self.h))
        if rc != 0:
            msg = self.lib.rekf_last_error(self.h).decode() if self.h else "creation failed"
            if self.h:
                self.lib.rekf_destroy(self.h)
            self.h = None
            raise RekfError(rc, msg)

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc):
        if rc < 0:
            raise RekfError(rc, self.lib.rekf_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.rekf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- hot path ---------------------------------------------------------------------------
    def handle_odometry(self, odom):
        """odom: (S, 4) float64 — time, vx, vy, wz per session."""
        odom = np.ascontiguousarray(odom, np.float64).reshape(self.S, 4)
        self._ck(self.lib.rekf_batch_handle_odometry(self.h, _ptr(odom)))

    def handle_observation(self, times, xy, counts=None):
        """times (S,), xy (S, m, 2) float32, counts (S,) or None (= m everywhere)."""
        xy = np.ascontiguousarray(xy, np.float32).reshape(self.S, -1, 2)
        m = xy.shape[1]
        times = np.ascontiguousarray(times, np.float64).reshape(self.S)
        counts = np.full(self.S, m, np.int32) if counts is None else np.ascontiguousarray(counts, np.int32).reshape(self.S)
        self._ck(self.lib.rekf_batch_handle_observation(self.h, _ptr(times), _ptr(xy), _ptr(counts), m))

    def handle_step(self, odom, times, xy, counts=None):
        """One whole step per session: the odometry message (S, 4) then the observation message — one call, one copy."""
        odom = np.ascontiguousarray(odom, np.float64).reshape(self.S, 4)
        xy = np.ascontiguousarray(xy, np.float32).reshape(self.S, -1, 2)
        m = xy.shape[1]
        times = np.ascontiguousarray(times, np.float64).reshape(self.S)
        counts = np.full(self.S, m, np.int32) if counts is None else np.ascontiguousarray(counts, np.int32).reshape(self.S)
        self._ck(self.lib.rekf_batch_handle_step(self.h, _ptr(odom), _ptr(times), _ptr(xy), _ptr(counts), m))

    def replay_device(self, d_odom, d_obs_time, d_obs_xy, T, m, d_pose_out=None):
        """Device pointers (ints): odom S x T x 4 f64, obs_time S x T f64, obs_xy S x T x m x 2 f32."""
        self._ck(self.lib.rekf_replay_device(self.h, d_odom, d_obs_time, d_obs_xy, T, m, d_pose_out))

    # -- accessors ----------------------------------------------------------------------------
    def sync(self):
        self._ck(self.lib.rekf_sync(self.h))

    def dim(self, s=0):
        return self._ck(self.lib.rekf_dim(self.h, s))

    def time(self, s=0):
        t = C.c_double()
        self._ck(self.lib.rekf_time(self.h, s, C.byref(t)))
        return t.value

    def mu(self, s=0):
        n = self.dim(s)
        out = np.zeros(n)
        nn = C.c_int()
        self._ck(self.lib.rekf_get_mu(self.h, s, _ptr(out), n, C.byref(nn)))
        return out

    def pose(self, s=0, with_cov=True):
        p = np.zeros(3)
        c = np.zeros((3, 3)) if with_cov else None
        self._ck(self.lib.rekf_get_pose(self.h, s, _ptr(p), _ptr(c)))
        return (p, c) if with_cov else p

    def poses(self, out=None):
        """(S, 3) poses of all sessions in one device→host read."""
        out = np.zeros((self.S, 3)) if out is None else out
        self._ck(self.lib.rekf_batch_get_pose(self.h, _ptr(out)))
        return out

    def request_poses(self):
        """Enqueue a device→host copy of all poses behind the work issued so far; returns a ticket (no host wait)."""
        t = C.c_int64()
        self._ck(self.lib.rekf_batch_request_poses(self.h, C.byref(t)))
        return t.value

    def fetch_poses(self, ticket, out=None):
        """Wait for `ticket` only and return its (S, 3) poses."""
        out = np.zeros((self.S, 3)) if out is None else out
        self._ck(self.lib.rekf_batch_fetch_poses(self.h, ticket, _ptr(out)))
        return out

    def landmarks(self, s=0):
        cnt = C.c_int()
        self._ck(self.lib.rekf_get_landmarks(self.h, s, None, None, 0, C.byref(cnt)))
        N = cnt.value
        xy = np.zeros((N, 2))
        cov = np.zeros((N, 2, 2))
        if N:
            self._ck(self.lib.rekf_get_landmarks(self.h, s, _ptr(xy), _ptr(cov), N, C.byref(cnt)))
        return xy, cov

    def markers(self, s=0):
        """(N, 5): x, y, angle, x_len, y_len of each landmark's 95 % covariance ellipse (ros_node.cc:736-789)."""
        cnt = C.c_int()
        self._ck(self.lib.rekf_get_markers(self.h, s, None, 0, C.byref(cnt)))
        out = np.zeros((cnt.value, 5))
        if cnt.value:
            self._ck(self.lib.rekf_get_markers(self.h, s, _ptr(out), cnt.value, C.byref(cnt)))
        return out

    def sigma(self, s=0):
        n = self.dim(s)
        out = np.zeros((n, n), order="F")
        self._ck(self.lib.rekf_get_sigma(self.h, s, out.ctypes.data_as(C.c_void_p), n))
        return np.ascontiguousarray(out)

    def state(self, s=0, n_expect=None, with_sigma=True):
        """(time, mu, sigma, flags) in one synchronisation (rekf_get_state)."""
        n = self.dim(s) if n_expect is None else n_expect
        mu = np.zeros(n)
        sig = np.zeros((n, n), order="F") if with_sigma else None
        t, nn, fl = C.c_double(), C.c_int(), C.c_int()
        self._ck(self.lib.rekf_get_state(self.h, s, n, C.byref(t), _ptr(mu), None if sig is None else sig.ctypes.data_as(C.c_void_p), n,
                                         C.byref(nn), C.byref(fl)))
        return t.value, mu, (None if sig is None else np.ascontiguousarray(sig)), fl.value

    def counters(self, s=0):
        """{updates, exact_frames, exact_slots}: how often the int8 covariance update left the tensor path (cumulative)."""
        out = (C.c_int64 * 3)()
        self._ck(self.lib.rekf_get_counters(self.h, s, out))
        return {"updates": out[0], "exact_frames": out[1], "exact_slots": out[2]}

    def match_result(self, s=0):
        cap = 1024
        sp = np.zeros((cap, 2), np.int32)
        mp = np.zeros((cap, 2), np.int32)
        nw = np.zeros(cap, np.int32)
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.rekf_get_match_result(self.h, s, _ptr(sp), C.byref(a), _ptr(mp), C.byref(b), _ptr(nw), C.byref(c), cap))
        return sp[: a.value].copy(), mp[: b.value].copy(), nw[: c.value].copy()

    def predict_state(self, time, s=0, with_sigma=True):
        n = self.dim(s)
        mu = np.zeros(n)
        sig = np.zeros((n, n), order="F") if with_sigma else None
        self._ck(self.lib.rekf_predict_state(self.h, s, time, _ptr(mu), n, None if sig is None else sig.ctypes.data_as(C.c_void_p), n))
        return mu, (None if sig is None else np.ascontiguousarray(sig))

    def set_state(self, time, vt, mu, sigma, s=0):
        mu = np.ascontiguousarray(mu, np.float64)
        n = mu.shape[0]
        sig = np.asfortranarray(np.asarray(sigma, np.float64))
        vt = np.ascontiguousarray(vt, np.float64)
        self._ck(self.lib.rekf_set_state(self.h, s, time, _ptr(vt), _ptr(mu), n, sig.ctypes.data_as(C.c_void_p), n))

    def set_map(self, xy, cov):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        cov = np.ascontiguousarray(cov, np.float64).reshape(-1, 4)
        self._ck(self.lib.rekf_set_map(self.h, _ptr(xy), _ptr(cov), xy.shape[0]))

    def get_map(self):
        cnt = C.c_int()
        self._ck(self.lib.rekf_get_map(self.h, None, None, 0, C.byref(cnt)))
        xy = np.zeros((cnt.value, 2), np.float32)
        cov = np.zeros((cnt.value, 4))
        if cnt.value:
            self._ck(self.lib.rekf_get_map(self.h, _ptr(xy), _ptr(cov), cnt.value, C.byref(cnt)))
        return xy, cov.reshape(-1, 2, 2)

    def load_map_txt(self, path):
        self._ck(self.lib.rekf_load_map_txt(self.h, path.encode()))

    def save_map_txt(self, filebase, s=0):
        self._ck(self.lib.rekf_save_map_txt(self.h, s, filebase.encode()))

    def debug_copy(self, name, count, dtype=np.float64, s=0):
        out = np.zeros(count, dtype)
        self._ck(self.lib.rekf_debug_copy(self.h, s, name.encode(), _ptr(out), out.nbytes))
        return out

    def error_flags(self, s=0):
        f = C.c_int()
        self._ck(self.lib.rekf_device_error_flags(self.h, s, C.byref(f)))
        return f.value

    # -- timing ---------------------------------------------------------------------------------
    def stream(self):
        return self.lib.rekf_stream(self.h)

    def timer_start(self):
        self._ck(self.lib.rekf_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.lib.rekf_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self._ck(self.lib.rekf_profile_enable(self.h, int(on)))

    def profile_read(self):
        cap = 16
        names = (C.c_char_p * cap)()
        us = (C.c_double * cap)()
        calls = (C.c_int * cap)()
        cnt = C.c_int()
        self._ck(self.lib.rekf_profile_read(self.h, names, us, calls, cap, C.byref(cnt)))
        return {names[k].decode(): (us[k], calls[k]) for k in range(cnt.value)}

    def launch_count(self):
        return self.lib.rekf_launch_count(self.h)


class OdometryData:
    """sensor::OdometryData (sensor/sensor_data.h:39-46) — only the fields the EKF reads (:216)."""

    def __init__(self, time, vx=0.0, vy=0.0, wz=0.0):
        self.time = time
        self.linear_velocity = (vx, vy, 0.0)
        self.angular_velocity = (0.0, 0.0, wz)


class Observation:
    """sensor::Observation (sensor/sensor_data.h:20-28): time_, cloud_ (float32 xy in base_link), gps_pose_."""

    def __init__(self, time, cloud, gps_pose=None):
        self.time_ = time
        self.cloud_ = np.ascontiguousarray(cloud, np.float32).reshape(-1, 2)
        self.gps_pose_ = gps_pose


class ReflectorEKFSLAM(EKFBatch):
    """Drop-in for ekf::ReflectorEKFSLAM (reflector_ekf_slam.h:13-64) over a single-session handle."""

    def __init__(self, options=None, **kw):
        super().__init__(1, options, **kw)

    def HandleOdometryMessage(self, odometry):                     # reflector_ekf_slam.cc:208-223
        self._ck(self.lib.rekf_handle_odometry(self.h, odometry.time, odometry.linear_velocity[0],
                                               odometry.linear_velocity[1], odometry.angular_velocity[2]))

    def HandleImuMessage(self, imu=None):                          # :224-227 (empty)
        self._ck(self.lib.rekf_handle_imu(self.h, 0.0))

    def HandleObservationMessage(self, observation):               # :229-368
        xy = observation.cloud_
        g = None if observation.gps_pose_ is None else np.ascontiguousarray(observation.gps_pose_, np.float64)
        self._ck(self.lib.rekf_handle_observation(self.h, observation.time_, _ptr(xy) if len(xy) else None, len(xy), _ptr(g)))

    def PredictState(self, time):                                  # :97-152
        mu, sig = self.predict_state(time)
        return time, mu, sig

    def GetStateVector(self):
        return self.mu(0)

    def GetCoviarance(self):
        return self.sigma(0)

    def GetLatestTime(self):
        return self.time(0)

    def GetState(self):
        return self.time(0), self.mu(0), self.sigma(0)

    def GetGlobalMap(self):
        return self.get_map()
