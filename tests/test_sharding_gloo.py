"""CPU: the N > 1 host logic (stream -> rank assignment, barrier, MAX/SUM reductions) with
torch.distributed gloo, world_size 2, on 127.0.0.1 - no GPU involved."""
import os
import socket

import pytest

from radiocapture_rf_b200 import sharding


def test_assignment_is_a_partition():
    for n, w in [(64, 8), (10, 4), (3, 8), (1, 1), (11, 2)]:
        seen = []
        for r in range(w):
            seen += sharding.assign_streams(n, w, r)
        assert sorted(seen) == list(range(n))
        per = sharding.streams_per_rank(n, w)
        assert max(per) - min(per) <= 1
    assert sharding.assign_streams(64, 8, 3) == [3, 11, 19, 27, 35, 43, 51, 59]   # BASELINE cfg 5: 8 per GPU
    with pytest.raises(ValueError):
        sharding.assign_streams(4, 2, 2)


def test_single_process_reducer_is_identity():
    r = sharding.Reducer()
    assert r.max(3.5) == 3.5 and r.sum(2) == 2.0
    rate, ms = sharding.whole_job_rate(1000, 2.0, r)
    assert rate == 500000.0 and ms == 2.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    red = sharding.Reducer(dist)
    mine = sharding.assign_streams(5, world, rank)            # rank0: 0,2,4  rank1: 1,3
    samples = 1000 * len(mine)
    ms = 10.0 + 5.0 * rank                                    # rank 1 is the slow one
    red.barrier()
    rate, worst = sharding.whole_job_rate(samples, ms, red)
    q.put((rank, mine, rate, worst, red.min(ms)))
    dist.destroy_process_group()


def test_two_rank_gloo_reduction():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    for rank, mine, rate, worst, best in res:
        assert worst == 15.0 and best == 10.0                 # MAX over ranks is what the bench reports
        assert abs(rate - 5000 / 15e-3) < 1e-6                # all ranks' units / slowest rank's time
