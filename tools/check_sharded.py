"""Multi-GPU check of the sharded belief (run under torchrun, one rank per GPU). Exit code 0 = all
ranks passed.

1. INVARIANTS. Every slot of every shard must hold a valid particle after each update — survivors,
   local duplicates and blocks stored by peers alike: its count block sums to prior + J * updates
   (J = FS + FO increments per update), its domain state is in range, its weight is 1/n, no cross-rank
   wait timed out; root-parallel rollouts return the same, bounded values on every rank.
2. SKEW. One rank's weights are multiplied by 10^3 before an update, so that rank owns ~99 % of all
   offspring and ships several times its own size to the other ranks (the old import buffer held
   n/64 records and silently dropped the rest). Its particles carry a distinctive domain state; the
   global posterior after the update must equal that of a SINGLE-GPU belief of world x n particles
   prepared the same way (number-of-set-bits histogram of the domain state, 5 sigma + 0.5 %).
3. POSTERIOR. Over several ordinary updates the global step likelihood and the state histogram of the
   sharded belief equal the single-GPU belief's within 5 standard errors (both are PHILOX-driven
   Monte-Carlo estimates of the same quantity; per-particle likelihood factors lie in [0, 1]).

usage: check_sharded.py [p2p|allgather] [n_local] [fixture]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

import fba_pomdp_b200 as fba
import golden_util as G
from fba_pomdp_b200.beliefs import _check
from fba_pomdp_b200.capi import ptr


def prototypes(g):
    sid, counts = g["is/init_struct_id"], g["is/init_counts"]
    seen, psid, pc = {}, [], []
    for i in range(len(sid)):
        k = (int(sid[i]), counts[i].tobytes())
        if k not in seen:
            seen[k] = len(psid)
            psid.append(int(sid[i]))
            pc.append(counts[i])
    return np.array(psid, np.int32), np.stack(pc)


def bits_hist(states, nbins):
    pop = np.array([bin(int(s)).count("1") for s in range(int(states.max()) + 1)])
    return np.bincount(pop[states], minlength=nbins)[:nbins].astype(np.float64)


def global_hist(local_states, nbins):
    h = torch.tensor(bits_hist(local_states, nbins), device="cuda", dtype=torch.float64)
    dist.all_reduce(h)
    return h.cpu().numpy()


def assert_same_distribution(h1, h2, what):
    n1, n2 = h1.sum(), h2.sum()
    p1, p2 = h1 / n1, h2 / n2
    p = (h1 + h2) / (n1 + n2)
    se = np.sqrt(p * (1 - p) * (1 / n1 + 1 / n2))
    bad = np.abs(p1 - p2) > 5.0 * se + 0.005
    assert not bad.any(), (what, p1, p2, se)


def set_particles(b, state=None, w=None):
    n = b.size()
    st = None if state is None else np.ascontiguousarray(state, np.int32)
    ww = None if w is None else np.ascontiguousarray(w, np.float64)
    _check(b.ctx.h, b.L.fba_belief_upload(b.h, 0, n, ptr(st), None, None, ptr(ww)))


def main():
    exchange = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    fixture = sys.argv[3] if len(sys.argv) > 3 else "sysadmin"
    journal = len(sys.argv) > 4 and sys.argv[4] == "journal"
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = G.load(fixture)
    ctx = fba.Context(local)
    J_nodes = len(np.asarray(g.desc["feat_s"]).reshape(-1)) + len(np.asarray(g.desc["feat_o"]).reshape(-1))
    desc = dict(g.desc, delta_capacity=J_nodes * 16) if journal else g.desc   # base + journal storage, 16 updates
    sim = fba.BAPOMDP(ctx, desc, g.t_par, g.o_par)
    psid, protos = prototypes(g)
    probs = None if len(psid) == 1 else np.ones(len(psid))
    J = sim.FS + sim.FO
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
    shared = np.random.RandomState(3)
    nbins = int(np.ceil(np.log2(sim.S))) + 1

    def make_sharded(seed):
        b = fba.ShardedBAImportanceSampling(n, exchange=exchange)
        rng = b.rank_rng(seed)
        b.initiate_sampled(sim, psid, protos, probs, rng, stride=protos.shape[1])
        return b, rng

    def make_single(seed):
        b = fba.BAImportanceSampling(n * world)
        rng = fba.Rng.philox(seed)
        b.initiate_sampled(sim, psid, protos, probs, rng, stride=protos.shape[1])
        return b, rng

    def check_slots(b, t_updates, base_sums):
        d = b.download()
        # a valid particle after t updates sums to a prototype's sum + J t; priors with fractional counts round
        # a little on every float32 +1, a stale or foreign block is off by at least one whole increment
        sums = d["counts"].astype(np.float64).sum(1) - float(J) * t_updates
        dist_to_proto = np.abs(sums[:, None] - base_sums[None, :]).min(1)
        assert dist_to_proto.max() < 0.25, (rank, t_updates, dist_to_proto.max(), np.unique(sums)[:5], base_sums[:5])
        assert d["state"].min() >= 0 and d["state"].max() < sim.S
        np.testing.assert_array_equal(d["w"], np.full(n, 1.0 / n))
        return d

    # the block sums the prior's prototypes have: a valid particle after t updates sums to one of them + J t
    base_sums = np.unique([float(c.astype(np.float64).sum()) for c in protos])

    # ---- 1. invariants + 3. posterior equivalence over ordinary updates ----
    b, rng = make_sharded(7)
    single = make_single(1007) if rank == 0 else None
    for t in range(8):
        a, o = script[t % len(script)]
        lik = b.updateEstimation(a, o, rng, step_uniform=float(shared.random_sample()))
        assert b.exchange == exchange, "fell back to %s" % b.exchange
        d = check_slots(b, t + 1, base_sums)
        h = global_hist(d["state"], nbins)
        if rank == 0:
            sb, srng = single
            slik = sb.updateEstimation(a, o, srng, want_likelihood=True)
            # shard weights are 1/n each, so the global total is a sum over `world` unit-mass shards
            se = 5.0 * np.sqrt(2.0) * 0.5 / np.sqrt(n * world)
            assert abs(lik / world - slik) <= se, (t, lik / world, slik, se)
            assert_same_distribution(h, bits_hist(sb.download(counts=False)["state"], nbins), "step %d" % t)
    assert b.timeouts() == 0
    if exchange == "allgather":
        assert b.L.fba_belief_dropped_records(b.h) == 0

    # root-parallel rollouts: 4099 requests split over the ranks, every rank gets all returns, in the
    # same order; bounded by the domain's reward range
    if fixture == "sysadmin":
        ret = b.rollouts(4099, 10, 0.95, rng)
        assert ret.shape == (4099,) and np.all(np.isfinite(ret))
        bound = 10.0 * (1 - 0.95 ** 10) / (1 - 0.95)   # |r| <= 10 (SysAdminBAExtension.cpp:27-48)
        assert ret.min() >= -bound - 1e-9 and ret.max() <= bound + 1e-9 and ret.mean() > 0, (ret.min(), ret.max())
        chk = torch.tensor([float(ret.sum()), float(ret[0]), float(ret[-1])], device="cuda", dtype=torch.float64)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on the gathered returns"
    b.free()
    if single:
        single[0].free()

    # ---- 2. skew: the last rank's weights x 1000, its particles in a distinctive state ----
    if exchange == "p2p":
        heavy = world - 1
        s_heavy, s_light = sim.S - 1, 0
        b, rng = make_sharded(11)
        set_particles(b, state=np.full(n, s_heavy if rank == heavy else s_light),
                      w=np.full(n, (1000.0 if rank == heavy else 1.0) / n))
        a, o = script[0]
        lik = b.updateEstimation(a, o, rng, step_uniform=0.37)
        d = check_slots(b, 1, base_sums)
        h = global_hist(d["state"], nbins)
        assert b.timeouts() == 0
        if rank == 0:
            sb, srng = make_single(2011)
            st = np.full(n * world, s_light, np.int32)
            w = np.full(n * world, 1.0 / (n * world))
            st[heavy * n:(heavy + 1) * n] = s_heavy
            w[heavy * n:(heavy + 1) * n] *= 1000.0
            set_particles(sb, state=st, w=w)
            sb.updateEstimation(a, o, srng)
            assert_same_distribution(h, bits_hist(sb.download(counts=False)["state"], nbins), "skew")
            sb.free()
        # and the belief keeps working afterwards
        for t in range(1, 4):
            a, o = script[t % len(script)]
            b.updateEstimation(a, o, rng, step_uniform=float(shared.random_sample()), likelihood=False)
        check_slots(b, 4, base_sums)
        assert b.timeouts() == 0
        b.free()

    sim.close()
    ctx.close()
    dist.barrier()
    if rank == 0:
        print("sharded check ok: exchange=%s world=%d n_local=%d fixture=%s%s"
              % (exchange, world, n, fixture, " (journal storage)" if journal else ""))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
