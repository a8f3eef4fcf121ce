"""Host-side multi-GPU logic on CPU: LPT bin partition and the QA-row all-gather over gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest


def test_partition_is_disjoint_complete_and_balanced():
    from checkm_b200.sharding import partition_bins
    rng = np.random.default_rng(0)
    costs = rng.lognormal(0, 0.8, size=1000)
    for world in (1, 2, 4, 8):
        parts = partition_bins(costs, world)
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(1000))
        loads = [costs[p].sum() for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.02
    assert [p.tolist() for p in partition_bins([5, 1, 1, 1, 1, 1], 2)] == [[0], [1, 2, 3, 4, 5]]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from checkm_b200.resultsParser import QA_DTYPE
    from checkm_b200.sharding import gather_rows, partition_bins
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    costs = [3.0, 1.0, 2.0, 2.5, 0.5]
    mine = partition_bins(costs, world)[rank]
    rows = np.zeros(len(mine), dtype=QA_DTYPE)
    for i, b in enumerate(mine):
        rows[i]['bin'] = b
        rows[i]['counts'] = [b, b + 1, 0, 0, 0, 0]
        rows[i]['completeness'] = 10.0 * b + 0.123456789
    allrows = gather_rows(rows, dist)
    q.put((rank, allrows.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_qa_rows_gloo_world2():
    import torch.multiprocessing as mp
    from checkm_b200.resultsParser import QA_DTYPE
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a = np.frombuffer(got[0], dtype=QA_DTYPE)
    b = np.frombuffer(got[1], dtype=QA_DTYPE)
    for name in QA_DTYPE.names:
        assert np.array_equal(a[name], b[name]), name
    assert sorted(a['bin'].tolist()) == [0, 1, 2, 3, 4]
    for r in a:
        assert r['completeness'] == 10.0 * r['bin'] + 0.123456789 and r['counts'][1] == r['bin'] + 1
