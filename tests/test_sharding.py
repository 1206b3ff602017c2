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


def test_bench_blocks_cover_the_job_once():
    """bench.py cuts the current x candidate grid of the job into one rectangular block per rank (shard_grid): every pair
    lands in exactly one block with the guess the unsharded job gives it, and a rank renders only its own frames"""
    sys.path.insert(0, ROOT)
    import bench
    assert [bench.shard_grid(64, 128, w) for w in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    n_cur, n_cand = 2, 4
    raws_cur, raws_cand, pairs, guesses = bench.make_workload(n_cur, n_cand, 3, procs=1)
    whole = {(int(r), int(c)): g for (r, c), g in zip(pairs, guesses)}
    assert len(whole) == n_cur * n_cand
    for world in (2, 4, 8):
        a, b = bench.shard_grid(n_cur, n_cand, world)
        assert a * b == world
        seen = {}
        for rank in range(world):
            rc, rk = rank // b, rank % b
            nc, nk = n_cur // a, n_cand // b
            rcur, rcand, p, g = bench.make_workload(n_cur, n_cand, 3, (rc * nc, (rc + 1) * nc), procs=1,
                                                    cand_slice=(rk * nk, (rk + 1) * nk))
            assert len(rcur) == nc and len(rcand) == nk and len(p) == nc * nk
            for (r, c), gg in zip(p, g):
                key = (int(r) + rk * nk, int(c) + rc * nc)
                assert key not in seen
                seen[key] = gg
                assert np.array_equal(gg, whole[key])
            # the frames are the job's frames
            for i in range(nc):
                assert np.array_equal(rcur[i], raws_cur[rc * nc + i])
            for i in range(nk):
                assert np.array_equal(rcand[i], raws_cand[rk * nk + i])
        assert len(seen) == n_cur * n_cand


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
