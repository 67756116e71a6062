"""CPU: the multi-GPU host logic (quota allocation, exchange plan, record exchange) — pure numpy
properties plus a world-size-2 run over gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fba_pomdp_b200 as fba


def test_quotas_sum_and_proportionality():
    rs = np.random.RandomState(0)
    for _ in range(200):
        G = rs.randint(1, 9)
        w = rs.gamma(2.0, size=G) + 1e-9
        n = int(rs.randint(1, 5000)) * G
        q = fba.offspring_quotas(w, n, rs.random_sample())
        assert q.sum() == n and (q >= 0).all()
        assert np.all(np.abs(q - n * w / w.sum()) <= 1.0 + 1e-9)  # systematic: within one of the share


def test_quotas_unbiased():
    w = np.array([0.2, 0.5, 0.3])
    rs = np.random.RandomState(1)
    acc = np.zeros(3)
    for _ in range(20000):
        acc += fba.offspring_quotas(w, 10, rs.random_sample())
    np.testing.assert_allclose(acc / 20000, 10 * w, atol=0.03)


def test_quotas_edge_cases():
    assert list(fba.offspring_quotas([1.0], 7, 0.3)) == [7]
    assert list(fba.offspring_quotas([0.0, 1.0], 6, 0.9)) == [0, 6]  # a shard whose weights collapsed
    with pytest.raises(fba.FbaError):
        fba.offspring_quotas([0.0, 0.0], 4, 0.1)


def test_exchange_plan_balances():
    rs = np.random.RandomState(2)
    for _ in range(200):
        G = rs.randint(1, 9)
        cap = int(rs.randint(1, 1000))
        w = rs.gamma(0.5, size=G) + 1e-12
        q = fba.offspring_quotas(w, cap * G, rs.random_sample())
        plan = fba.exchange_plan(q, cap)
        assert (np.diag(plan) == 0).all() and (plan >= 0).all()
        after = q - plan.sum(1) + plan.sum(0)
        assert (after == cap).all()
        # only over-quota ranks send, only under-quota ranks receive
        assert (plan.sum(1)[q <= cap] == 0).all() and (plan.sum(0)[q >= cap] == 0).all()
    with pytest.raises(fba.FbaError):
        fba.exchange_plan([3, 3], 2)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # all-gather of the shard totals -> identical quotas and plan on every rank
        local = [3.0, 1.0][rank]
        mine = torch.tensor([local], dtype=torch.float64)
        tot = torch.empty(world, dtype=torch.float64)
        dist.all_gather_into_tensor(tot, mine)
        cap, rb = 8, 24
        q = fba.offspring_quotas(tot.numpy(), cap * world, 0.25)
        plan = fba.exchange_plan(q, cap)
        # records: byte r of a record shipped by rank g carries (g, r)
        n_out = int(plan[rank].sum())
        src = torch.zeros(n_out * rb, dtype=torch.uint8)
        for r in range(n_out):
            src[r * rb:(r + 1) * rb] = 16 * rank + r
        got = fba.exchange_records(dist, None, plan, rank, src, rb)
        # root-parallel rollouts: 7 requests over 2 ranks -> 4 + 3, returns gathered in rank order
        counts = fba.split_requests(7, world)
        mine_ret = 100.0 * rank + np.arange(counts[rank], dtype=np.float64)
        allret = fba.gather_ragged(dist, None, mine_ret, counts)
        out.put((rank, q.tolist(), plan.tolist(), got.numpy().tolist(), counts.tolist(), allret.tolist()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, q0, plan0, got0, c0, r0), (_, q1, plan1, got1, c1, r1) = res
    assert c0 == c1 == [4, 3]
    assert r0 == r1 == [0.0, 1.0, 2.0, 3.0, 100.0, 101.0, 102.0]
    assert q0 == q1 == [12, 4] and plan0 == plan1 == [[0, 4], [0, 0]]
    assert got0 == []  # rank 0 is over quota: receives nothing
    assert len(got1) == 4 * 24 and got1[::24] == [0, 1, 2, 3]  # rank 1 received rank 0's 4 records


def test_python_and_library_plans_agree():
    """offspring_quotas / exchange_plan (numpy, used by the gloo tests) restate what
    fba_belief_shard_resample computes in C; the GPU test test_shard_plan_matches_python checks the
    C side against them. Here: the numpy pair is self-consistent on the edge the C code clamps."""
    q = fba.offspring_quotas([1e-300, 1.0, 1e-300], 30, 0.999)
    assert q.sum() == 30 and q[1] >= 29


def test_split_requests_properties():
    for n in (0, 1, 7, 4096, 4099):
        for g in (1, 2, 3, 8):
            c = fba.split_requests(n, g)
            assert c.sum() == n and c.max() - c.min() <= 1 and list(c) == sorted(c, reverse=True)
    np.testing.assert_array_equal(fba.gather_ragged(None, None, [1.5, 2.5], [2]), [1.5, 2.5])
