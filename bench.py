#!/usr/bin/env python
"""bench.py — EKF steps/s at N = 1024 landmarks / 100 observed per step (BASELINE.json config C3).

One "step" = one HandleOdometryMessage + one HandleObservationMessage (2 predicts + association + update)
for every session of the batch.  Per GPU the workload is `--sessions` independent sessions advancing
through the same launches (BASELINE config 5 = 64 C3 sessions over 8 GPUs = 8 per GPU; that per-GPU slice is
the default at every N, so scaling is weak).

  value          sessions x steps / device time, inputs resident in HBM (rekf_replay_device; the CUDA graphs are
                 captured during warm-up, the timed window only launches them)
  e2e            the same metric through the host-buffer C-ABI calls (rekf_batch_handle_step + the poses of every step
                 read back), >= 100 steps whatever --steps says; e2e.blocking = host waits for the pose after every step
  e2e_adapter    the drop-in path: the C++ adapter driven like the reference node (Handle* + GetState() by value after
                 every message, one session), and the same with GetPose() only
  single_session one session alone (latency-bound figure, Sigma L2-resident)
  parity         the timed run itself checked against the CPU oracle (structured algebra, pinned to the reference's own
                 code by tests/test_reference_pin.py): pose trajectory, final mu / Sigma, association lists
  roofline       the covariance GEMM alone (events around that launch only)
  cpu_baseline   the reference's OWN translation unit (oracle/_ref) on one host core — what its build produces;
                 cpu_baseline_structured / cpu_baseline_blas: context, so that dropping the reference's redundant n^3
                 algebra is not credited to the GPU

    python bench.py [--gpus N] [--steps K] [--warmup W] [--sessions S] [--config C2|C3|C4] [--impl reference]

For N > 1 launch under torchrun (one rank per GPU); ranks never communicate inside a step — NCCL only
scatters the synthetic input streams from rank 0, gathers the results and reduces the timing (max over ranks).
"""
import argparse
import json
import os
import struct
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "steps/s"
# name -> (N landmarks, m observed per step, odometry model, default sessions per GPU)
CONFIGS = {"C2": (256, 50, "diff", 8), "C3": (1024, 100, "diff", 8), "C4": (4096, 200, "omni", 1)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--sessions", type=int, default=0, help="independent sessions per GPU (0: the configuration's default)")
    ap.add_argument("--groups", type=int, default=1, help="pipeline groups the sessions of one GPU are split into (1: lock-step)")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs the persistent SYRK leaves to other groups (0: engine default)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cov", default="i8", choices=["i8", "tcgen05", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip every CPU leg (cpu_baseline*, parity)")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--parity-steps", type=int, default=40, help="steps of the timed run the CPU oracle follows (all of them if W+K is smaller)")
    ap.add_argument("--no-adapter", action="store_true", help="skip the C++ adapter leg")
    a = ap.parse_args()
    a.N, a.m, a.model, dflt = CONFIGS[a.config]
    if a.sessions <= 0:
        a.sessions = dflt
    a.warmup = max(a.warmup, 3)
    return a


def metric_name(a):
    return f"EKF steps/sec at N={a.N} landmarks ({a.m} observed/step)"


def config_dict(a, world):
    """The workload description — the SAME object in both arms (the reference arm runs this workload on host cores)."""
    S, G = a.sessions, max(1, min(a.groups, a.sessions))
    n_int = 4 + 2 * a.N
    ld = -(-n_int // 128) * 128
    sig_mb = ld * ld * 8 / 1e6
    return {
        "workload": f"{a.config}: synthetic 2D stream, N={a.N} landmarks, {a.m} observed/step, {a.model} odom; "
                    f"{S} independent sessions per GPU" + (" (BASELINE config 5 per-GPU slice)" if a.config == "C3" and S == 8 else ""),
        "sessions_per_gpu": S, "pipeline_groups": G, "n": 3 + 2 * a.N, "r": 2 * a.m, "cov_update": a.cov,
        "l2": f"working set {S} x {sig_mb:.0f} MB Sigma buffers ({S * sig_mb / 2:.0f} MB of upper triangles touched per step) "
              + ("> 126 MB L2 (inputs larger than L2)" if S * sig_mb / 2 > 126 else "<= 126 MB L2: L2-resident between steps, stated here"),
    }


# ---------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"          # B200_PROFILING.md fallback figures


def syrk_algorithmic_bytes(n, r):
    """SURVEY.md §8(d) bold row with Sigma in fp64 (w = 8): read Sigma + write Sigma + read W^T once."""
    return 2.0 * n * n * 8 + r * n * 8


def syrk_implemented_bytes(n, r):
    """What the kernel has to move per session: the upper triangle of Sigma read and written once (by the L2, through the
    TMA reduce-add) + the four int8 digit panels of W^T read once (K padded to 64)."""
    return 2.0 * (n * n / 2) * 8 + 4.0 * n * (-(-r // 64) * 64)


def build_streams(a, sessions, steps_total, session_offset=0):
    from reflector_ekf_slam_b200.synth import make_stream
    return [make_stream(a.config, steps_total, session=session_offset + s) for s in range(sessions)]


def warm_start(batch, streams, N):
    """Map-building phase through the engine's own augmentation path (untimed)."""
    nb = streams[0]["n_build"]
    for k in range(nb):
        batch.handle_odometry(np.stack([st["odom"][k] for st in streams]))
        batch.handle_observation(np.array([st["obs_time"][k] for st in streams]), np.stack([st["obs_xy"][k] for st in streams]),
                                 np.array([st["obs_count"][k] for st in streams]))
    batch.sync()
    for s in range(len(streams)):
        assert batch.dim(s) == 3 + 2 * N, f"session {s}: map building produced n = {batch.dim(s)}"
    return nb


# ---------------------------------------------------------------------------------------------------
# CPU legs (the checker side: oracle/ — used here ONLY as the measured CPU baseline and the parity checker)
# ---------------------------------------------------------------------------------------------------
def cpu_snapshot(a, st):
    """State after the map-building phase, from the structured-algebra oracle (seconds; same numbers as the as-written
    algebra to 1e-12, tests/test_reference_pin.py)."""
    from oracle.pyoracle import STRUCTURED, Oracle
    builder = Oracle(algebra=STRUCTURED, native=True, odom_model=0 if a.model == "diff" else 1)
    for k in range(st["n_build"]):
        builder.HandleOdometryMessage(*st["odom"][k])
        builder.HandleObservationMessage(st["obs_time"][k], st["obs_xy"][k, : st["obs_count"][k]])
    t, mu, sig = builder.GetState()
    return t, st["odom"][st["n_build"] - 1][1:4], mu, sig


def make_cpu_filter(a, kind):
    """kind: 'reference' = the reference's own reflector_ekf_slam.cc compiled into oracle/_ref (falls back to the C port
    where that library is absent), 'structured' = C port without the exact-zero work."""
    from oracle import pyoracle
    from oracle.pyoracle import AS_WRITTEN, STRUCTURED, Oracle
    model = 0 if a.model == "diff" else 1
    if kind == "structured":
        return Oracle(algebra=STRUCTURED, native=True, odom_model=model), "port"
    try:
        return pyoracle.Reference(fast=True, track_matches=False, odom_model=model), "reference"
    except Exception:
        return Oracle(algebra=AS_WRITTEN, native=True, odom_model=model), "port"


def time_cpu_steps(a, kind, threads, warmup, steps, budget_s):
    """`threads` independent sessions, one per host thread (the reference is single-threaded, CMakeLists.txt:4-6), each
    advancing `warmup` untimed + up to `steps` timed steps of the same stream from the map-building snapshot.  All threads
    run the same number of steps; the count shrinks so that the whole call fits `budget_s`.
    Returns dict(value, steps, warmup, seconds, kind, s_per_step)."""
    from oracle import pyoracle
    pyoracle.build(native=True)
    st = build_streams(a, 1, warmup + steps + 1)[0]
    snap = cpu_snapshot(a, st)
    filters, kinds = zip(*[make_cpu_filter(a, kind) for _ in range(threads)])
    for f in filters:
        f.set_state(*snap)
    nb = st["n_build"]
    plan = {"W": warmup, "K": steps}
    bar = threading.Barrier(threads + 1)
    t_first = [0.0] * threads

    def advance(f, k):
        f.HandleOdometryMessage(*st["odom"][nb + k])
        f.HandleObservationMessage(st["obs_time"][nb + k], st["obs_xy"][nb + k])

    def worker(i):
        f = filters[i]
        t0 = time.perf_counter()
        advance(f, 0)                                   # first warm-up step, timed to size the rest
        t_first[i] = time.perf_counter() - t0
        bar.wait()
        bar.wait()                                      # the main thread fixed W and K
        for k in range(1, plan["W"]):
            advance(f, k)
        bar.wait()
        for k in range(plan["W"], plan["W"] + plan["K"]):
            advance(f, k)
        bar.wait()

    ths = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(threads)]
    for t in ths:
        t.start()
    bar.wait()
    t1 = max(t_first)
    afford = max(1, int(budget_s / max(t1, 1e-6)) - 1)          # steps that still fit after the first one
    plan["W"] = max(1, min(warmup, 1 + afford // 4))
    plan["K"] = max(1, min(steps, afford - (plan["W"] - 1)))
    bar.wait()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    seconds = time.perf_counter() - t0
    for t in ths:
        t.join()
    return {"value": threads * plan["K"] / seconds, "steps": plan["K"], "warmup": plan["W"], "seconds": seconds, "kind": kinds[0],
            "s_per_step": seconds / plan["K"]}


def time_blas_steps(a, steps, budget_s):
    """Context: the reference's as-written dense products through numpy / OpenBLAS on every host core (oracle/numpy_ekf.py)."""
    from oracle.numpy_ekf import NumpyEKF
    st = build_streams(a, 1, steps + 1)[0]
    t, vt, mu, sig = cpu_snapshot(a, st)
    f = NumpyEKF(odom_model=0 if a.model == "diff" else 1)
    f.time, f.vt, f.mu, f.sigma = t, np.array(vt, float), mu.copy(), sig.copy()
    nb, done, t0 = st["n_build"], 0, time.perf_counter()
    while done < steps:
        f.handle_odometry(*st["odom"][nb + done])
        f.handle_observation(st["obs_time"][nb + done], st["obs_xy"][nb + done])
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    sec = time.perf_counter() - t0
    return {"value": done / sec, "steps": done, "seconds": sec}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(a):
    """The reference arm: the reference's own CPU implementation of the path (oracle/_ref: reflector_ekf_slam.cc compiled
    unmodified, -O3 -mavx2 -mfma build where the CPU has it) on every host core, one independent session per thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    budget = 150.0 if a.cpu_budget_s >= 25.0 else max(30.0, 4 * a.cpu_budget_s)
    res = time_cpu_steps(a, "reference", cores, a.warmup, a.steps, budget_s=budget)
    desc = (f"{a.config}: {'the reference translation unit (oracle/_ref, unmodified reflector_ekf_slam.cc over the Eigen stand-in)' if res['kind'] == 'reference' else 'as-written dense fp64 C port (oracle/_ref absent)'}, "
            f"{cores} host thread(s) x 1 session each, {res['warmup']} warm-up + {res['steps']} timed step(s) per thread, "
            f"{res['s_per_step']:.2f} s per step per thread; CPU: {cpu_model()}")
    out = {
        "impl": "reference", "metric": metric_name(a), "value": res["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": res["steps"],
        "warmup": res["warmup"], "ms_per_step": 1e3 * res["seconds"] / res["steps"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(a, a.gpus),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": cores, "kind": res["kind"], "sample": desc,
                         "per_core_steps_s": res["value"] / cores},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "requested": {"steps": a.steps, "warmup": a.warmup}, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------
# parity of the timed run (CPU oracle in a background thread while the GPU legs run)
# ---------------------------------------------------------------------------------------------------
class ParityChecker:
    """Follows one session's stream on the CPU (structured algebra) for the first P steps after map building and keeps the
    pose after each step, the association lists and the final state."""

    def __init__(self, a, stream, steps):
        self.a, self.st, self.P = a, stream, steps
        self.poses, self.final, self.err = None, None, None
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        try:
            from oracle import pyoracle
            pyoracle.build(native=True)
            f, _ = make_cpu_filter(self.a, "structured")
            st, nb = self.st, self.st["n_build"]
            for k in range(nb):
                f.HandleOdometryMessage(*st["odom"][k])
                f.HandleObservationMessage(st["obs_time"][k], st["obs_xy"][k, : st["obs_count"][k]])
            poses = np.zeros((self.P, 3))
            for k in range(self.P):
                f.HandleOdometryMessage(*st["odom"][nb + k])
                f.HandleObservationMessage(st["obs_time"][nb + k], st["obs_xy"][nb + k])
                poses[k] = f.GetStateVector()[:3]
            self.poses = poses
            self.final = (f.GetStateVector(), f.GetCoviarance(), f.match_result())
        except Exception as e:                              # reported in the JSON line, never fatal for the timing
            self.err = repr(e)

    def result(self, gpu_poses, gpu_final):
        """gpu_poses: (>= P, 3) poses after each step from the start of warm-up; gpu_final: (mu, Sigma, match lists) after
        exactly P steps, or None when the GPU ran further than the oracle followed."""
        self.th.join()
        if self.err:
            return {"checked": False, "error": self.err}
        d = np.abs(gpu_poses[: self.P] - self.poses)
        d[:, 2] = np.abs((d[:, 2] + np.pi) % (2 * np.pi) - np.pi)
        out = {"checked": True, "against": "oracle (structured fp64 algebra; == the reference's own code to 1e-12, tests/test_reference_pin.py)",
               "steps_compared": int(self.P), "pose_xy_max_m": float(d[:, :2].max()), "pose_yaw_max_rad": float(d[:, 2].max()),
               "tolerance": {"dmu_m": 1e-4, "sigma_rel_fro": 1e-5}}
        if gpu_final is not None:
            mu, sig, matches = gpu_final
            omu, osig, omatches = self.final
            out["dmu"] = float(np.abs(mu - omu).max()) if mu.shape == omu.shape else None
            out["relfro"] = float(np.linalg.norm(sig - osig) / np.linalg.norm(osig)) if sig.shape == osig.shape else None
            out["matches_equal"] = bool(all(np.array_equal(x, y) for x, y in zip(matches, omatches)))
            out["ok"] = bool(out["dmu"] is not None and out["dmu"] < 1e-4 and out["relfro"] < 1e-5 and out["matches_equal"] and out["pose_xy_max_m"] < 1e-4)
        else:
            out["ok"] = bool(out["pose_xy_max_m"] < 1e-4)
            out["note"] = "final mu / Sigma not compared: the timed run is longer than --parity-steps"
        return out


# ---------------------------------------------------------------------------------------------------
# the drop-in path: C++ adapter, node call pattern
# ---------------------------------------------------------------------------------------------------
def adapter_leg(a, stream, k_state, k_pose):
    lib_dir = os.path.join(ROOT, "reflector_ekf_slam_b200")
    exe = os.path.join(ROOT, "scripts", "bin", "adapter_bench")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(ROOT, "scripts", "adapter_bench.cc")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        cmd = ["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "stubs"), src, "-o", exe,
               "-L", lib_dir, "-l:librekf_b200.so", f"-Wl,-rpath,{lib_dir}"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            return {"error": "g++ failed: " + res.stderr[-300:]}
    nb = stream["n_build"]
    T = nb + 2 * k_state + k_pose + 2
    path = os.path.join(ROOT, "scripts", "bin", f"adapter_stream_{os.getpid()}.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("4i", T, a.m, a.N, 0 if a.model == "diff" else 1))
        for k in range(T):
            f.write(stream["odom"][k].astype(np.float64).tobytes())
            f.write(struct.pack("d", float(stream["obs_time"][k])))
            f.write(struct.pack("i", int(stream["obs_count"][k])))
            f.write(np.ascontiguousarray(stream["obs_xy"][k], np.float32).tobytes())
    try:
        res = subprocess.run([exe, path, str(nb), str(k_state), str(k_pose)], capture_output=True, text=True, timeout=300)
    finally:
        os.unlink(path)
    if res.returncode != 0:
        return {"error": f"adapter_bench rc={res.returncode}: {res.stderr[-300:]}"}
    d = json.loads(res.stdout.strip().splitlines()[-1])
    n = d["n"]
    return {
        "value": d["state_steps"] / d["state_seconds"], "unit": UNIT, "steps": d["state_steps"], "sessions": 1,
        "api": "ekf::ReflectorEKFSLAMB200 (C++11 adapter): HandleOdometryMessage, GetState(), HandleObservationMessage, GetState() per step — "
               "the node's pattern (ros_node.cc:515,638): a by-value State with the full n x n covariance after every message",
        "d2h_bytes_per_step": 2 * (n * n + n + 3) * 8,
        "by_reference": {"value": d["state_steps"] / d["ref_seconds"], "steps": d["state_steps"], "d2h_bytes_per_step": 2 * (n * n + n + 3) * 8,
                         "api": "same, GetStateVector() + GetCoviarance() (references to the refreshed, page-locked mirror: no by-value copy)"},
        "pose_only": {"value": d["pose_steps"] / d["pose_seconds"], "steps": d["pose_steps"], "d2h_bytes_per_step": 2 * 12 * 8,
                      "api": "same, GetPose() (pose + 3x3 block) in place of GetState()"},
    }


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    from reflector_ekf_slam_b200 import build as rbuild
    from reflector_ekf_slam_b200.engine import EKFBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if rank == 0:
        rbuild.build()
    if world > 1:
        dist.barrier()

    S, K, W, N_LM, M_OBS = a.sessions, a.steps, a.warmup, a.N, a.m
    T = W + K
    KP = min(K, 50)                            # per-kernel profiling pass
    KE = 100 if a.config != "C4" else 20       # end-to-end pass: >= 100 steps whatever --steps says (C4: 20, a step is milliseconds)
    EXTRA = KP + 2 * KE
    cov = {"tcgen05": 0, "f64": 1, "i8": 2}[a.cov]
    odom_model = 0 if a.model == "diff" else 1

    # ---- inputs: rank 0 generates every session's stream, NCCL scatters the packed shards ----------
    nb = int(np.ceil(N_LM / M_OBS))
    if world > 1:
        Ttot = nb + T + EXTRA
        from reflector_ekf_slam_b200.shard import scatter_streams
        streams = scatter_streams(lambda: build_streams(a, S * world, T + EXTRA), S, (Ttot, 6 + 2 * M_OBS), nb, dev)
    else:
        streams = build_streams(a, S, T + EXTRA)

    # the CPU oracle follows this rank's session 0 through the first P steps of warm-up + timed run, in the background
    P = min(T, a.parity_steps)
    checker = None if a.no_cpu_baseline else ParityChecker(a, dict(streams[0], n_build=nb), P)

    G = max(1, min(a.groups, S))
    batch = EKFBatch(S, max_landmarks=N_LM, max_observations=M_OBS, device=local, cov_update=cov, use_graphs=1,
                     pipeline_groups=G, syrk_reserve_sms=a.reserve_sms, odom_model=odom_model)
    warm_start(batch, streams, N_LM)

    def dev_inputs(lo, hi):
        d_odom = torch.tensor(np.stack([st["odom"][lo:hi] for st in streams]), device=dev)
        d_time = torch.tensor(np.stack([st["obs_time"][lo:hi] for st in streams]), device=dev)
        d_xy = torch.tensor(np.stack([st["obs_xy"][lo:hi] for st in streams]), device=dev)
        return d_odom, d_time, d_xy

    # ---- (1) device-resident replay: W warm-up steps (graphs captured here), then exactly K timed steps ----
    w_in = dev_inputs(nb, nb + W)
    k_in = dev_inputs(nb + W, nb + W + K)
    d_pose_w = torch.zeros(S, W, 3, dtype=torch.float64, device=dev)
    d_pose = torch.zeros(S, K, 3, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    Pw = min(P, W)                             # if the oracle stops inside warm-up, snapshot the engine there
    gpu_final = None
    if 0 < P <= W:
        batch.replay_device(w_in[0].data_ptr(), w_in[1].data_ptr(), w_in[2].data_ptr(), P, M_OBS, d_pose_w.data_ptr())
        batch.sync()
        gpu_final = (batch.mu(0), batch.sigma(0), batch.match_result(0)) if checker else None
        if W > P:
            rest = dev_inputs(nb + P, nb + W)
            batch.replay_device(rest[0].data_ptr(), rest[1].data_ptr(), rest[2].data_ptr(), W - P, M_OBS, None)
    else:
        batch.replay_device(w_in[0].data_ptr(), w_in[1].data_ptr(), w_in[2].data_ptr(), W, M_OBS, d_pose_w.data_ptr())
    batch.sync()
    cnt0 = [batch.counters(s) for s in range(S)]
    launches0 = batch.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    batch.timer_start()
    batch.replay_device(k_in[0].data_ptr(), k_in[1].data_ptr(), k_in[2].data_ptr(), K, M_OBS, d_pose.data_ptr())
    ms = batch.timer_stop()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    clocks = sampler.stop() if rank == 0 else None
    launches = batch.launch_count() - launches0
    batch.sync()                                       # raises if any session flagged an error
    cnt1 = [batch.counters(s) for s in range(S)]
    exact = {k: int(sum(c1[k] - c0[k] for c0, c1 in zip(cnt0, cnt1))) for k in ("updates", "exact_frames", "exact_slots")}
    sp, _, nw = batch.match_result(0)
    steady = (len(sp) == M_OBS and len(nw) == 0 and batch.dim(0) == 3 + 2 * N_LM)
    value = world * S * K / (ms * 1e-3)
    if checker and P == T:                             # the oracle followed the whole timed run: compare the final state too
        gpu_final = (batch.mu(0), batch.sigma(0), batch.match_result(0))
    mu_end, sig_end = batch.mu(0), None
    pose_traj = torch.cat([d_pose_w, d_pose], 1).cpu().numpy()          # (S, W+K, 3)

    # ---- (2) per-kernel device times over the same kind of steps (events around every launch) -----------
    p_in = dev_inputs(nb + W + K, nb + W + K + KP)
    batch.profile_enable(True)
    batch.replay_device(p_in[0].data_ptr(), p_in[1].data_ptr(), p_in[2].data_ptr(), KP, M_OBS, None)
    prof = batch.profile_read()
    batch.profile_enable(False)
    step_us = sum(v[0] for v in prof.values())
    syrk_name = {0: "k_syrk_tcgen05", 1: "k_syrk_f64", 2: "k_syrk_tcgen05_i8"}[cov]
    syrk_us = prof[syrk_name][0]
    n_ref, r = 3 + 2 * N_LM, 2 * M_OBS
    hbm_peak, bf16_peak, peak_kind = measured_peaks()
    Sl = -(-S // G)                                    # sessions per launch: every kernel is launched once per pipeline group
    alg_bytes = syrk_algorithmic_bytes(n_ref, r) * Sl
    impl_bytes = syrk_implemented_bytes(n_ref, r) * Sl
    achieved = alg_bytes / (syrk_us * 1e-6) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "syrk_traffic.json")
    if os.path.exists(tp) and a.config == "C3":
        try:
            tj = json.load(open(tp))
            traffic, traffic_src = tj.get(f"S{Sl}"), "constant from profiles/syrk_traffic.json (one ncu --set full capture of this kernel at this launch size; not measured in this run): " + tj.get("source", "")
        except Exception:
            traffic = None
    flops_useful = 2.0 * n_ref * n_ref * r * Sl
    tensor = {"useful_tflops": flops_useful / (syrk_us * 1e-6) / 1e12}
    if cov == 2:
        # upper triangle only, ten 128x64x32 int8 products (digit pairs p+q <= 3) per useful fp64-class product
        tensor.update({"executed_int8_tops": 10 * 0.5 * flops_useful * (-(-r // 32) * 32 / r) / (syrk_us * 1e-6) / 1e12,
                       "int8_peak_tops": 2 * bf16_peak,
                       "note": "kind::i8 tcgen05: 10 integer products per useful product over the upper triangle, K padded to 32; int8 peak taken as twice the measured bf16 peak"})
    elif cov == 0:
        tensor.update({"executed_tflops": 3 * 0.5 * flops_useful / (syrk_us * 1e-6) / 1e12, "tf32_peak_tflops": bf16_peak / 2,
                       "note": "3 tf32 products per useful product over the upper triangle; tf32 peak taken as half the measured bf16 peak"})
    roofline = {
        "kernel": syrk_name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": alg_bytes, "sessions_per_launch": Sl, "kernel_us": syrk_us, "share_of_step": syrk_us / step_us,
        "implemented_bytes_per_launch": impl_bytes, "frac_of_implemented_bytes": impl_bytes / (syrk_us * 1e-6) / 1e9 / hbm_peak,
        "note": "achieved = SURVEY 8(d) algorithmic bytes (full Sigma read + written, fp64) / kernel time; the kernel itself keeps only the upper triangle of Sigma "
                "and lets the L2 do the read-modify-write (TMA reduce-add), so it moves about half of that (implemented_bytes): frac can exceed 1",
        "tensor": tensor, "kernels_us": {k: round(v[0], 2) for k, v in prof.items()},
    }

    # ---- (3) end to end through the host-buffer C ABI: H2D of every message, D2H of the poses, every step --
    lo = nb + W + K + KP
    od = np.stack([st["odom"][lo:lo + 2 * KE] for st in streams], 1).copy()          # (2 KE, S, 4)
    ot = np.stack([st["obs_time"][lo:lo + 2 * KE] for st in streams], 1).copy()
    ox = np.stack([st["obs_xy"][lo:lo + 2 * KE] for st in streams], 1).copy()        # (2 KE, S, m, 2)
    poses = np.zeros((S, 3))
    batch.sync()
    if world > 1:
        dist.barrier()
    # (3a) blocking form: the pose is read back (and waited for) after every step, like the node's GetState()
    t0 = time.perf_counter()
    for k in range(KE):
        batch.handle_step(od[k], ot[k], ox[k])
        batch.poses(poses)
    e2e_block_s = time.perf_counter() - t0
    # (3b) streaming form: every step's poses still come back to the host, through the pinned ring, but the host
    # redeems a step's ticket LAG steps later, so the groups keep their stagger
    LAG = 8
    pending = []
    t0 = time.perf_counter()
    for k in range(KE, 2 * KE):
        batch.handle_step(od[k], ot[k], ox[k])
        pending.append(batch.request_poses())
        if len(pending) > LAG:
            batch.fetch_poses(pending.pop(0), poses)
    for t in pending:
        batch.fetch_poses(t, poses)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([e2e_s, e2e_block_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s, e2e_block_s = float(tmax[0].item()), float(tmax[1].item())
    h2d = S * (4 * 8) + S * (8 + 4 * 8 + 4) + S * M_OBS * 8
    e2e = {"value": world * S * KE / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": S * 24, "steps": KE,
           "api": "rekf_batch_handle_step (odometry + observation message of every session, host buffers) + rekf_batch_request_poses per step ("
                  f"poses of every step out through the pinned ring, ticket redeemed {LAG} steps later)",
           "blocking": {"value": world * S * KE / e2e_block_s, "steps": KE,
                        "api": "same calls with rekf_batch_get_pose (host waits for the pose) after every step"}}
    batch.sync()

    # ---- (4) multi-GPU evidence: gather of results; the same seeded session on every rank must be bit-identical ----
    multi = None
    if world > 1:
        from reflector_ekf_slam_b200.synth import make_stream
        V = 12
        vs = make_stream(a.config, V, session=0)               # the SAME session on every rank
        one = EKFBatch(1, max_landmarks=N_LM, max_observations=M_OBS, device=local, cov_update=cov, use_graphs=1, odom_model=odom_model)
        warm_start(one, [vs], N_LM)
        for k in range(nb, nb + V):
            one.handle_step(vs["odom"][k][None], vs["obs_time"][k:k + 1], vs["obs_xy"][k][None])
        one.sync()
        import zlib
        digest = zlib.crc32(one.mu(0).tobytes(), zlib.crc32(np.ascontiguousarray(one.sigma(0)).tobytes()))
        one.close()
        mine = torch.tensor([float(digest), float(zlib.crc32(np.ascontiguousarray(pose_traj).tobytes())),
                             float(np.abs(mu_end).sum())], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        traj = torch.tensor(pose_traj[:, -1, :], device=dev)                         # final pose of every session of this rank
        trajs = [torch.zeros_like(traj) for _ in range(world)]
        dist.all_gather(trajs, traj)
        digests = [int(v[0].item()) for v in allv]
        multi = {"same_seed_session_bit_identical_across_ranks": bool(len(set(digests)) == 1), "state_crc32_per_rank": digests,
                 "pose_trajectory_crc32_per_rank": [int(v[1].item()) for v in allv],
                 "final_poses_gathered": [[round(float(x), 6) for x in t[0].tolist()] for t in trajs],
                 "note": f"every rank also ran global session 0 for {V} steps after map building: CRC32 of (mu, Sigma) must agree bit for bit; "
                         "pose trajectories (S x (W+K) x 3) and final states of the timed sessions are gathered to rank 0"}

    # ---- (5) parity of the timed run against the CPU oracle ------------------------------------------------
    parity = None
    if checker:
        parity = checker.result(pose_traj[0], gpu_final)
        if world > 1:
            ok = torch.tensor([1.0 if parity.get("ok") else 0.0, parity.get("pose_xy_max_m", 0.0) or 0.0], dtype=torch.float64, device=dev)
            oks = [torch.zeros_like(ok) for _ in range(world)]
            dist.all_gather(oks, ok)
            parity["ok_all_ranks"] = bool(all(o[0].item() == 1.0 for o in oks))
            parity["pose_xy_max_m_all_ranks"] = float(max(o[1].item() for o in oks))

    # ---- (6) one session alone (latency-bound figure), rank 0 ----------------------------------------------
    single = None
    if rank == 0:
        one = EKFBatch(1, max_landmarks=N_LM, max_observations=M_OBS, device=local, cov_update=cov, use_graphs=1, odom_model=odom_model)
        warm_start(one, streams[:1], N_LM)
        Ks = min(K, 200)
        b1 = [torch.tensor(streams[0][key][nb:nb + W][None], device=dev) for key in ("odom", "obs_time", "obs_xy")]
        one.replay_device(b1[0].data_ptr(), b1[1].data_ptr(), b1[2].data_ptr(), W, M_OBS, None)
        one.sync()
        b2 = [torch.tensor(streams[0][key][nb + W:nb + W + Ks][None], device=dev) for key in ("odom", "obs_time", "obs_xy")]
        one.timer_start()
        one.replay_device(b2[0].data_ptr(), b2[1].data_ptr(), b2[2].data_ptr(), Ks, M_OBS, None)
        ms1 = one.timer_stop()
        single = {"value": Ks / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1 / Ks, "steps": Ks,
                  "note": "one session per GPU: the upper triangle of Sigma stays L2-resident between steps"}
        one.close()

    # ---- (7) the drop-in path and the CPU baselines (rank 0, N = 1 only) --------------------------------------
    adapter = cpu = cpu_struct = cpu_blas = None
    if rank == 0 and world == 1:
        batch.close()
        batch = None
        if not a.no_adapter:
            try:
                adapter = adapter_leg(a, streams[0], 30 if a.config != "C4" else 4, 140 if a.config != "C4" else 20)
            except Exception as e:
                adapter = {"error": repr(e)}
        if not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            r1 = time_cpu_steps(a, "reference", 1, 1, 5, a.cpu_budget_s)
            cpu = {"value": r1["value"], "unit": UNIT, "cores": 1, "kind": r1["kind"],
                   "sample": f"{a.config}: {'the reference translation unit (oracle/_ref)' if r1['kind'] == 'reference' else 'as-written dense fp64 C port'}, "
                             f"1 thread (the reference build is single-threaded, CMakeLists.txt:4-6), {r1['warmup']} warm-up + {r1['steps']} timed step(s), "
                             f"{r1['s_per_step']:.2f} s/step; {cores} host cores present; CPU: {cpu_model()}"}
            r2 = time_cpu_steps(a, "structured", 1, 1, 20, 10.0)
            cpu_struct = {"value": r2["value"], "unit": UNIT, "cores": 1, "kind": "port",
                          "sample": f"{a.config}: structure-exploiting fp64 C variant of the same equations (no exact-zero work: O(n) predict, block-sparse H, "
                                    f"Cholesky + rank-r downdate), 1 thread, {r2['steps']} timed step(s), {r2['s_per_step'] * 1e3:.0f} ms/step — context: the algebraic saving, not GPU speed"}
            try:
                r3 = time_blas_steps(a, 5, 15.0)
                cpu_blas = {"value": r3["value"], "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{a.config}: the as-written dense products through numpy/OpenBLAS on all {cores} host threads (oracle/numpy_ekf.py), {r3['steps']} step(s)"}
            except Exception as e:
                cpu_blas = {"error": repr(e)}

    if rank == 0:
        out = {
            "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f64 state/solve + tf32x3 tcgen05 covariance GEMM (fp32 TMEM accumulate)", 1: "f64",
                      2: "f64 state/solve + exact int8-slice (4x7-bit) tcgen05 covariance GEMM (s32 TMEM accumulate)"}[cov],
            "data": "synthetic", "config": dict(config_dict(a, world), steady_state_all_matched=bool(steady)),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "cpu_baseline_structured": cpu_struct, "cpu_baseline_blas": cpu_blas, "e2e_adapter": adapter, "single_session": single,
            "parity": parity, "int8_exact_path": dict(exact, note="timed window, all sessions of this rank: updates = frames with an update; exact_frames = "
                                                                  "whole frames rerouted to the fp64 SYRK; exact_slots = flagged slots redone in fp64 by k_syrk_exact_rows"),
        }
        if multi:
            out["multi_gpu"] = multi
        print(json.dumps(out), flush=True)
    if batch is not None:
        batch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
