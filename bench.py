#!/usr/bin/env python
"""bench.py — EKF steps/s at N = 1024 landmarks / 100 observed per step (BASELINE.json config C3).

One "step" = one HandleOdometryMessage + one HandleObservationMessage (2 predicts + association + update)
for every session of the batch.  Per GPU the workload is `--sessions` independent C3 sessions advancing
through the same launches (BASELINE config 5 = 64 sessions over 8 GPUs = 8 per GPU; that per-GPU slice is
the default at every N, so scaling is weak).  `value` = sessions x steps / device time, inputs resident in
HBM (rekf_replay_device).  `e2e` = the same metric through the host-buffer C-ABI calls
(rekf_batch_handle_step + the poses of every step read back).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--sessions S] [--impl reference]

For N > 1 launch under torchrun (one rank per GPU); ranks never communicate inside a step — NCCL only
scatters the synthetic input streams from rank 0 and reduces the timing (max over ranks).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "EKF steps/sec at N=1024 landmarks (100 observed/step)"
UNIT = "steps/s"
CONFIG = "C3"
N_LM, M_OBS = 1024, 100


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--sessions", type=int, default=8, help="independent sessions per GPU")
    ap.add_argument("--groups", type=int, default=2, help="pipeline groups the sessions of one GPU are split into (1: lock-step)")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs the persistent SYRK leaves to other groups (0: engine default)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cov", default="i8", choices=["i8", "tcgen05", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--single-session", action="store_true", help="also time one session alone (latency-bound figure)")
    return ap.parse_args()


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
    """SURVEY.md §8(d) bold row with Σ in fp64 (w = 8): read Σ + write Σ + read both tf32 panels of Wᵀ."""
    return 2.0 * n * n * 8 + r * n * 8


def build_streams(sessions, steps_total, session_offset=0):
    from reflector_ekf_slam_b200.synth import make_stream
    return [make_stream(CONFIG, steps_total, session=session_offset + s) for s in range(sessions)]


def warm_start(batch, streams):
    """Map-building phase through the engine's own augmentation path (untimed)."""
    nb = streams[0]["n_build"]
    for k in range(nb):
        batch.handle_odometry(np.stack([st["odom"][k] for st in streams]))
        batch.handle_observation(np.array([st["obs_time"][k] for st in streams]), np.stack([st["obs_xy"][k] for st in streams]),
                                 np.array([st["obs_count"][k] for st in streams]))
    batch.sync()
    for s in range(len(streams)):
        assert batch.dim(s) == 3 + 2 * N_LM, f"session {s}: map building produced n = {batch.dim(s)}"
    return nb


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algebra (as-written dense fp64 restatement, oracle/) on host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference(budget_s, max_steps, threads, native=True):
    """Time the as-written oracle at C3 on `threads` host threads (one independent session per thread — the
    reference itself is single-threaded, CMakeLists.txt:4-6).  Returns (steps/s aggregate, description)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    from oracle.pyoracle import AS_WRITTEN, STRUCTURED, Oracle
    try:
        pyoracle.build(native=native)
    except Exception:
        native = False
    streams = build_streams(1, max_steps + 1)
    st = streams[0]
    builder = Oracle(algebra=STRUCTURED, native=native)       # untimed warm start (structured algebra, same numbers)
    for k in range(st["n_build"]):
        builder.HandleOdometryMessage(*st["odom"][k])
        builder.HandleObservationMessage(st["obs_time"][k], st["obs_xy"][k, : st["obs_count"][k]])
    t, mu, sig = builder.GetState()
    vt = st["odom"][st["n_build"] - 1][1:4]

    def run(_):
        o = Oracle(algebra=AS_WRITTEN, native=native)
        o.set_state(t, vt, mu, sig)
        done, t0 = 0, time.perf_counter()
        k = st["n_build"]
        while done < max_steps:
            o.HandleOdometryMessage(*st["odom"][k + done])
            o.HandleObservationMessage(st["obs_time"][k + done], st["obs_xy"][k + done])
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        return done, time.perf_counter() - t0

    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(run, range(threads)))
    total_steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    per_thread = [r[1] / r[0] for r in res]
    desc = (f"C3 as-written dense fp64 oracle ({'-O3 -march=native' if native else '-O3'}), {threads} thread(s) x "
            f"{res[0][0]} step(s), median {np.median(per_thread):.2f} s/step/thread")
    return total_steps / wall, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    value, desc = cpu_reference(budget_s=max(20.0, args.cpu_budget_s * 4), max_steps=max(1, min(args.steps, 3)), threads=cores)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * cores / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{CONFIG}: N={N_LM} landmarks, {M_OBS} observed/step, one CPU session per host thread"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args):
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

    S, K, W = args.sessions, args.steps, args.warmup
    T = W + K
    EXTRA = min(K, 50) + min(K, 100)          # profiling pass + end-to-end pass
    cov = {"tcgen05": 0, "f64": 1, "i8": 2}[args.cov]

    # ---- inputs: rank 0 generates every session's stream, NCCL scatters the packed shards ----------
    streams = None
    if world > 1:
        nb = int(np.ceil(N_LM / M_OBS))
        Ttot = nb + T + EXTRA
        from reflector_ekf_slam_b200.shard import scatter_streams
        streams = scatter_streams(lambda: build_streams(S * world, T + EXTRA), S, (Ttot, 6 + 2 * M_OBS), nb, dev)
    else:
        streams = build_streams(S, T + EXTRA)

    G = max(1, min(args.groups, S))
    batch = EKFBatch(S, max_landmarks=N_LM, max_observations=M_OBS, device=local, cov_update=cov, use_graphs=1,
                     pipeline_groups=G, syrk_reserve_sms=args.reserve_sms)
    nb = warm_start(batch, streams)

    def dev_inputs(lo, hi):
        d_odom = torch.tensor(np.stack([st["odom"][lo:hi] for st in streams]), device=dev)
        d_time = torch.tensor(np.stack([st["obs_time"][lo:hi] for st in streams]), device=dev)
        d_xy = torch.tensor(np.stack([st["obs_xy"][lo:hi] for st in streams]), device=dev)
        return d_odom, d_time, d_xy

    # ---- (1) device-resident replay: W warm-up steps, then exactly K timed steps ------------------------
    w_in = dev_inputs(nb, nb + W)
    k_in = dev_inputs(nb + W, nb + W + K)
    d_pose = torch.zeros(S, K, 3, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    if W > 0:
        batch.replay_device(w_in[0].data_ptr(), w_in[1].data_ptr(), w_in[2].data_ptr(), W, M_OBS, None)
    batch.sync()
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
    sp, _, nw = batch.match_result(0)
    steady = (len(sp) == M_OBS and len(nw) == 0 and batch.dim(0) == 3 + 2 * N_LM)
    value = world * S * K / (ms * 1e-3)

    # ---- (2) per-kernel device times over the same kind of steps (events around every launch) -----------
    p_in = dev_inputs(nb + W + K, nb + W + K + min(K, 50))
    batch.profile_enable(True)
    batch.replay_device(p_in[0].data_ptr(), p_in[1].data_ptr(), p_in[2].data_ptr(), min(K, 50), M_OBS, None)
    prof = batch.profile_read()
    batch.profile_enable(False)
    step_us = sum(v[0] for v in prof.values())
    syrk_name = {0: "k_syrk_tcgen05", 1: "k_syrk_f64", 2: "k_syrk_tcgen05_i8"}[cov]
    syrk_us = prof[syrk_name][0]
    n_ref, r = 3 + 2 * N_LM, 2 * M_OBS
    hbm_peak, bf16_peak, peak_kind = measured_peaks()
    Sl = -(-S // G)                                    # sessions per launch: every kernel is launched once per pipeline group
    alg_bytes = syrk_algorithmic_bytes(n_ref, r) * Sl
    achieved = alg_bytes / (syrk_us * 1e-6) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "syrk_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"S{Sl}")
        except Exception:
            traffic = None
    flops_useful = 2.0 * n_ref * n_ref * r * Sl
    roofline = {
        "kernel": syrk_name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "traffic": traffic,
        "algorithmic_bytes_per_launch": alg_bytes, "sessions_per_launch": Sl, "kernel_us": syrk_us, "share_of_step": syrk_us / step_us,
        "tensor": {"useful_tflops": flops_useful / (syrk_us * 1e-6) / 1e12, "executed_tflops": 0.75 * flops_useful / (syrk_us * 1e-6) / 1e12,
                   "tf32_peak_tflops": bf16_peak / 2, "note": "tf32 peak taken as half the measured bf16 peak; executed = 3 tf32 products over the upper triangle only"},
        "kernels_us": {k: round(v[0], 2) for k, v in prof.items()},
    }

    # ---- (3) end to end through the host-buffer C ABI: H2D of every message, D2H of the poses, every step --
    lo = nb + W + K + min(K, 50)
    Ke = min(K, 100)
    od = np.stack([st["odom"][lo:lo + Ke] for st in streams], 1).copy()          # (Ke, S, 4)
    ot = np.stack([st["obs_time"][lo:lo + Ke] for st in streams], 1).copy()
    ox = np.stack([st["obs_xy"][lo:lo + Ke] for st in streams], 1).copy()        # (Ke, S, m, 2)
    poses = np.zeros((S, 3))
    batch.sync()
    if world > 1:
        dist.barrier()
    # (3a) blocking form: the pose is read back (and waited for) after every step, like the node's GetState()
    Kb = Ke // 2
    t0 = time.perf_counter()
    for k in range(Kb):
        batch.handle_step(od[k], ot[k], ox[k])
        batch.poses(poses)
    e2e_block_s = time.perf_counter() - t0
    # (3b) streaming form: every step's poses still come back to the host, through the pinned ring, but the host
    # redeems a step's ticket LAG steps later, so the groups keep their stagger
    LAG = 8
    pending = []
    t0 = time.perf_counter()
    for k in range(Kb, Ke):
        batch.handle_step(od[k], ot[k], ox[k])
        pending.append(batch.request_poses())
        if len(pending) > LAG:
            batch.fetch_poses(pending.pop(0), poses)
    for t in pending:
        batch.fetch_poses(t, poses)
    e2e_s = time.perf_counter() - t0
    Ke_stream = Ke - Kb
    if world > 1:
        tmax = torch.tensor([e2e_s, e2e_block_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s, e2e_block_s = float(tmax[0].item()), float(tmax[1].item())
    e2e_value = world * S * Ke_stream / e2e_s
    h2d = S * (4 * 8) + S * (8 + 4 * 8 + 4) + S * M_OBS * 8
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": S * 24, "steps": Ke_stream,
           "api": "rekf_batch_handle_step (odometry + observation message of every session, host buffers) + rekf_batch_request_poses per step ("
                  f"poses of every step out through the pinned ring, ticket redeemed {LAG} steps later)",
           "blocking": {"value": world * S * Kb / e2e_block_s, "steps": Kb,
                        "api": "same calls with rekf_batch_get_pose (host waits for the pose) after every step"}}

    # ---- (4) optional: one session alone (latency-bound single-stream figure) ----------------------------
    single = None
    if args.single_session and rank == 0:
        one = EKFBatch(1, max_landmarks=N_LM, max_observations=M_OBS, device=local, cov_update=cov, use_graphs=1)
        warm_start(one, streams[:1])
        a = [torch.tensor(streams[0][key][nb:nb + T][None], device=dev) for key in ("odom", "obs_time", "obs_xy")]
        one.replay_device(a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), W, M_OBS, None)
        one.sync()
        b = [torch.tensor(streams[0][key][nb + W:nb + W + K][None], device=dev) for key in ("odom", "obs_time", "obs_xy")]
        one.timer_start()
        one.replay_device(b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), K, M_OBS, None)
        ms1 = one.timer_stop()
        single = {"value": K / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1 / K, "note": "Sigma (34 MB) stays L2-resident between steps"}
        one.close()

    # ---- (5) CPU baseline (rank 0, N = 1 only): the reference's as-written algebra on the host -----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, desc = cpu_reference(budget_s=args.cpu_budget_s, max_steps=5, threads=1)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f64 state/solve + tf32x3 tcgen05 covariance GEMM (fp32 TMEM accumulate)", 1: "f64",
                      2: "f64 state/solve + exact int8-slice (4x7-bit) tcgen05 covariance GEMM (s32 TMEM accumulate)"}[cov],
            "data": "synthetic",
            "config": {"workload": f"{CONFIG}: synthetic 2D stream, N={N_LM} landmarks, {M_OBS} observed/step, diff odom; "
                                   f"{S} independent sessions per GPU (BASELINE config 5 per-GPU slice) in {G} pipeline group(s)",
                       "sessions_per_gpu": S, "pipeline_groups": G, "n": n_ref, "r": r, "cov_update": args.cov,
                       "l2": f"working set {S} x 38 MB Sigma = {S * 38} MB per step " + ("> 126 MB L2 (inputs larger than L2)" if S * 38 > 126 else "<= L2: see single_session note"),
                       "steady_state_all_matched": bool(steady)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        if single:
            out["single_session"] = single
        print(json.dumps(out), flush=True)
    batch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
