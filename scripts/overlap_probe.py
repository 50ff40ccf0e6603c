"""Does a second handle on its own stream fill the SMs the latency-bound front half leaves idle?
usage: overlap_probe.py [total sessions] [handles] [steps]"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from reflector_ekf_slam_b200.engine import EKFBatch
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
W = 5
dev = torch.device("cuda:0")
streams = bench.build_streams(S, W + K)
per = S // G
hs, ins = [], []
for g in range(G):
    b = EKFBatch(per, max_landmarks=bench.N_LM, max_observations=bench.M_OBS, cov_update=2, use_graphs=1)
    sub = streams[g * per:(g + 1) * per]
    nb = bench.warm_start(b, sub)
    def dev_inputs(lo, hi, sub=sub):
        return [torch.tensor(np.stack([st[k][lo:hi] for st in sub]), device=dev) for k in ("odom", "obs_time", "obs_xy")]
    w = dev_inputs(nb, nb + W); k = dev_inputs(nb + W, nb + W + K)
    b.replay_device(w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(), W, bench.M_OBS, None)
    hs.append(b); ins.append(k)
for b in hs: b.sync()
torch.cuda.synchronize()
t0 = time.perf_counter()
for b, k in zip(hs, ins):
    b.replay_device(k[0].data_ptr(), k[1].data_ptr(), k[2].data_ptr(), K, bench.M_OBS, None)
for b in hs: b.sync()
dt = time.perf_counter() - t0
print(f"S={S} handles={G} steps={K}: {S * K / dt:.0f} steps/s  ({dt / K * 1e6:.1f} us per lock-step of {S})")
