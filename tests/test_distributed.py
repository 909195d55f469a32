"""N>1 host logic on CPU: two gloo ranks partition the shots and reduce their counters."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out):
    import torch.distributed as dist
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank), "MASTER_ADDR": "127.0.0.1",
                       "MASTER_PORT": str(port)})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quits_b200.distributed import my_shots, reduce_results
    lo, hi = my_shots(total)
    # stand-in for the per-rank Monte-Carlo result: counters that depend only on the global shot indices
    idx = np.arange(lo, hi, dtype=np.int64)
    counts = np.array([np.sum(idx % 7 == 0), np.sum(idx % 11 == 0), hi - lo], dtype=np.int64)
    c, t = reduce_results(counts, [10.0 + rank, 5.0 - rank])
    out.put((rank, lo, hi, c.tolist(), t.tolist()))
    dist.destroy_process_group()


def test_two_ranks_partition_and_reduce():
    total, world = 100_003, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, c0, t0), (r1, lo1, hi1, c1, t1) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 == total and lo1 % 64 == 0
    idx = np.arange(total)
    want = [int(np.sum(idx % 7 == 0)), int(np.sum(idx % 11 == 0)), total]
    assert c0 == want and c1 == want
    assert t0 == [11.0, 5.0] and t1 == [11.0, 5.0]
