"""N > 1 path on CPU: world_size-2 gloo run of the session scatter (the only collective on the path) — every rank
must receive exactly its shard of the seeded streams (checksummed), and ranks given the same seed must agree."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reflector_ekf_slam_b200.shard import pack_streams, scatter_streams, unpack_streams
from reflector_ekf_slam_b200.synth import make_stream, stream_checksum


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, per_rank, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = make_stream("T1", steps, session=0)
    T, m, nb = ref["obs_xy"].shape[0], ref["m"], ref["n_build"]
    mine = scatter_streams(lambda: [make_stream("T1", steps, session=s) for s in range(world * per_rank)],
                           per_rank, (T, 6 + 2 * m), nb, torch.device("cpu"))
    sums = [stream_checksum(st) for st in mine]
    expect = [stream_checksum(make_stream("T1", steps, session=rank * per_rank + s)) for s in range(per_rank)]
    ok = sums == expect
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out.put((rank, ok, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_pack_unpack_roundtrip_is_exact():
    sts = [make_stream("T1", 5, session=s) for s in range(3)]
    back = unpack_streams(pack_streams(sts), sts[0]["n_build"])
    for a, b in zip(sts, back):
        assert np.array_equal(a["odom"], b["odom"]) and np.array_equal(a["obs_time"], b["obs_time"])
        assert np.array_equal(a["obs_xy"], b["obs_xy"]) and np.array_equal(a["obs_count"], b["obs_count"])


def test_world2_gloo_scatter_delivers_each_ranks_shard():
    world, per_rank = 2, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_rank, 4, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [2.0, 2.0]
