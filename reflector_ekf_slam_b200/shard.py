"""Session sharding across ranks (multi-GPU = independent replay sessions; SURVEY.md §8e: replicas only).

The EKF path never communicates inside a step.  The only collective is the one-off scatter of the packed
synthetic input streams from rank 0 (and the max-over-ranks reduction of the timing in bench.py).  Works with
any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def pack_streams(streams):
    """(sessions, T, 6 + 2m) float64: odom(4) | obs_time | obs_count | obs_xy (float32 values, exactly representable)."""
    T, m = streams[0]["obs_xy"].shape[0], streams[0]["obs_xy"].shape[1]
    out = np.zeros((len(streams), T, 6 + 2 * m))
    for s, st in enumerate(streams):
        out[s, :, 0:4] = st["odom"]
        out[s, :, 4] = st["obs_time"]
        out[s, :, 5] = st["obs_count"]
        out[s, :, 6:] = st["obs_xy"].reshape(T, -1)
    return out


def unpack_streams(packed, n_build):
    packed = np.asarray(packed)
    S, T, w = packed.shape
    m = (w - 6) // 2
    return [{"odom": packed[s, :, 0:4].copy(), "obs_time": packed[s, :, 4].copy(),
             "obs_count": packed[s, :, 5].astype(np.int32), "obs_xy": packed[s, :, 6:].astype(np.float32).reshape(T, m, 2),
             "n_build": n_build, "m": m} for s in range(S)]


def scatter_streams(make_all, sessions_per_rank, shape_tail, n_build, device):
    """Rank 0 calls make_all() -> list of world*sessions_per_rank streams; every rank gets its shard back."""
    world, rank = dist.get_world_size(), dist.get_rank()
    recv = torch.empty((sessions_per_rank,) + tuple(shape_tail), dtype=torch.float64, device=device)
    if rank == 0:
        streams = make_all()
        assert len(streams) == world * sessions_per_rank
        chunks = [torch.tensor(pack_streams(streams[g * sessions_per_rank:(g + 1) * sessions_per_rank]), device=device)
                  for g in range(world)]
        dist.scatter(recv, chunks, src=0)
    else:
        dist.scatter(recv, None, src=0)
    return unpack_streams(recv.cpu().numpy(), n_build)
