"""N>1 host logic on CPU: world_size-2 gloo run of the pair partition + record gather."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from g2o_frontend_b200 import capi, sharding


def test_partition_covers_everything():
    for n in (0, 1, 7, 8192, 1000):
        for w in (1, 2, 3, 4, 8):
            blocks = [sharding.partition(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_order_pairs_groups_currents():
    pairs = np.array([[3, 1], [0, 0], [2, 1], [5, 0], [1, 2]])
    perm = sharding.order_pairs_by_current(pairs)
    cur = pairs[perm][:, 1]
    assert list(cur) == sorted(cur)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.partition(n_total, world, rank)
    rec = np.zeros(hi - lo, capi.RESULT_DTYPE)
    rec["inliers"] = np.arange(lo, hi)
    rec["T"][:, 0] = np.arange(lo, hi) * 0.5
    rec["status"] = rank
    allrec = sharding.gather_records(rec, n_total)
    q.put((rank, allrec["inliers"].tolist(), allrec["status"].tolist(), allrec["T"][:, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 16])
def test_gather_records_gloo_world2(n_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, inl, status, t0 in outs:
        assert inl == list(range(n_total))           # complete and ordered
        lo1, hi1 = sharding.partition(n_total, 2, 1)
        assert status == [0] * lo1 + [1] * (hi1 - lo1)
        assert t0 == [i * 0.5 for i in range(n_total)]
