"""ctypes binding of the CPU oracle (oracle/rekf_oracle.c).

TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  PARITY UNPINNED by the reference (see rekf_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from reflector_ekf_slam_b200._abi import RekfOptions, make_options  # noqa: F401  (re-export)

_HERE = os.path.dirname(os.path.abspath(__file__))
AS_WRITTEN = 0
STRUCTURED = 1
# oracle/_ref only: instantiate ekf::ReflectorEKFSLAMGPS (reflector_ekf_slam_gps.cc) instead of ekf::ReflectorEKFSLAM
REF_GPS_CLASS = 0x100

REFERENCE_ROOT = "/root/reference"
REF_DIR = os.path.join(_HERE, "_ref")

_libs = {}


def build(native=False):
    """(Re)build liboracle.so (and liboracle_native.so) with the committed Makefile."""
    target = "native" if native else "all"
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True)


def build_ref():
    """Compile the reference's own translation units into oracle/_ref/ (only where /root/reference exists:
    this container; the GPU box uses the prebuilt files that travel with the snapshot)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return False
    subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)
    return True


def ref_available():
    return os.path.exists(os.path.join(REF_DIR, "librekf_ref.so")) or os.path.isdir(REFERENCE_ROOT)


def _cpu_has_avx2_fma():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " avx2" in flags and " fma" in flags


def load_ref(fast=False):
    """oracle/_ref/librekf_ref.so: ekf::ReflectorEKFSLAM{,GPS} compiled unmodified from /root/reference
    against oracle/shim (see ref_abi.cc), same ABI as liboracle.so.  fast=True picks the -mavx2 -mfma build
    when the CPU has it (timing only; the reference's own CMake build is the plain -O3 one)."""
    name = "librekf_ref_avx2.so" if (fast and _cpu_has_avx2_fma()) else "librekf_ref.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(REF_DIR, name)
    if os.path.isdir(REFERENCE_ROOT):
        build_ref()          # make: no-op when up to date
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: build it with `make -C oracle ref` where /root/reference exists")
    lib = C.CDLL(path)
    _bind(lib, ref=True)
    _libs[name] = lib
    return lib


def load(native=False):
    name = "liboracle_native.so" if native else "liboracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "rekf_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build(native)
    lib = C.CDLL(path)
    _bind(lib, ref=False)
    _libs[name] = lib
    return lib


def _bind(lib, ref):
    P = C.POINTER
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_create.argtypes = [P(RekfOptions), C.c_int]
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_handle_odometry.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.oracle_handle_observation.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
    lib.oracle_dim.argtypes = [C.c_void_p]
    lib.oracle_dim.restype = C.c_int
    lib.oracle_time.argtypes = [C.c_void_p]
    lib.oracle_time.restype = C.c_double
    lib.oracle_mu.argtypes = [C.c_void_p]
    lib.oracle_mu.restype = P(C.c_double)
    lib.oracle_sigma.argtypes = [C.c_void_p]
    lib.oracle_sigma.restype = P(C.c_double)
    lib.oracle_get_match_result.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int]
    lib.oracle_predict_state.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    lib.oracle_set_state.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.oracle_set_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.oracle_get_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.oracle_get_map.restype = C.c_int
    lib.oracle_load_map_txt.argtypes = [C.c_void_p, C.c_char_p]
    lib.oracle_save_map_txt.argtypes = [C.c_void_p, C.c_char_p]
    lib.oracle_save_map_txt.restype = C.c_int
    if ref:
        lib.oracle_ref_track_matches.argtypes = [C.c_void_p, C.c_int]
    else:
        lib.oracle_dgemm.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.oracle_lu_inverse.argtypes = [C.c_int, C.c_void_p, C.c_int]
        lib.oracle_lu_inverse.restype = C.c_int


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """One CPU filter.  Method names follow the reference class (reflector_ekf_slam.h:20-44)."""

    def __init__(self, options=None, algebra=STRUCTURED, native=False, lib=None, **kw):
        self.lib = lib if lib is not None else load(native)
        self.options = options if options is not None else make_options(**kw)
        self.h = self.lib.oracle_create(C.byref(self.options), int(algebra))
        self.algebra = algebra

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- hot path -----------------------------------------------------------------------
    def HandleOdometryMessage(self, time, vx, vy, wz):
        self.lib.oracle_handle_odometry(self.h, time, vx, vy, wz)

    def HandleObservationMessage(self, time, xy, gps_pose=None):
        xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
        g = None if gps_pose is None else np.ascontiguousarray(gps_pose, dtype=np.float64)
        self.lib.oracle_handle_observation(self.h, time, _ptr(xy), xy.shape[0], _ptr(g))

    # --- accessors ----------------------------------------------------------------------
    def dim(self):
        return self.lib.oracle_dim(self.h)

    def GetLatestTime(self):
        return self.lib.oracle_time(self.h)

    def GetStateVector(self):
        n = self.dim()
        return np.ctypeslib.as_array(self.lib.oracle_mu(self.h), shape=(n,)).copy()

    def GetCoviarance(self):
        n = self.dim()
        flat = np.ctypeslib.as_array(self.lib.oracle_sigma(self.h), shape=(n * n,)).copy()
        return flat.reshape(n, n).T.copy()  # column-major storage → [row, col]

    def GetState(self):
        return self.GetLatestTime(), self.GetStateVector(), self.GetCoviarance()

    def match_result(self):
        cap = 4096
        sp = np.zeros((cap, 2), np.int32)
        mp = np.zeros((cap, 2), np.int32)
        nw = np.zeros(cap, np.int32)
        ns, nm, nn = C.c_int(), C.c_int(), C.c_int()
        self.lib.oracle_get_match_result(self.h, _ptr(sp), C.addressof(ns), _ptr(mp), C.addressof(nm),
                                         _ptr(nw), C.addressof(nn), cap)
        return sp[: ns.value].copy(), mp[: nm.value].copy(), nw[: nn.value].copy()

    def PredictState(self, time):
        n = self.dim()
        mu = np.zeros(n)
        sig = np.zeros(n * n)
        self.lib.oracle_predict_state(self.h, time, _ptr(mu), _ptr(sig))
        return mu, sig.reshape(n, n).T.copy()

    def set_state(self, time, vt, mu, sigma):
        mu = np.ascontiguousarray(mu, np.float64)
        n = mu.shape[0]
        sig = np.asfortranarray(np.asarray(sigma, np.float64))
        vt = np.ascontiguousarray(vt, np.float64)
        self.lib.oracle_set_state(self.h, time, _ptr(vt), _ptr(mu), n, sig.ctypes.data_as(C.c_void_p), n)

    def set_map(self, xy, cov):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        cov = np.ascontiguousarray(cov, np.float64).reshape(-1, 4)
        self.lib.oracle_set_map(self.h, _ptr(xy), _ptr(cov), xy.shape[0])

    def GetGlobalMap(self):
        cap = 65536
        xy = np.zeros((cap, 2), np.float32)
        cov = np.zeros((cap, 4), np.float64)
        n = self.lib.oracle_get_map(self.h, _ptr(xy), _ptr(cov), cap)
        return xy[:n].copy(), cov[:n].reshape(-1, 2, 2).copy()

    def load_map_txt(self, path):
        self.lib.oracle_load_map_txt(self.h, path.encode())

    def save_map_txt(self, filebase):
        return self.lib.oracle_save_map_txt(self.h, filebase.encode())


class Reference(Oracle):
    """The reference's own ekf::ReflectorEKFSLAM (gps=True: ekf::ReflectorEKFSLAMGPS), compiled unmodified
    from /root/reference into oracle/_ref/ — the pin for the restatement above and for the CUDA engine."""

    def __init__(self, options=None, gps=False, fast=False, track_matches=True, **kw):
        super().__init__(options, algebra=(REF_GPS_CLASS if gps else 0), lib=load_ref(fast), **kw)
        if not track_matches:
            self.lib.oracle_ref_track_matches(self.h, 0)
