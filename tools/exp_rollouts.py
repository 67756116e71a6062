"""Scratch experiment: rollout kernel variants at small batch (not part of the product)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fba_pomdp_b200 as fba
import golden_util as G
g = G.load("gridworld3")
n = int(os.environ.get("N", 200000))
ctx = fba.Context(0)
sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
b = fba.BAImportanceSampling(n)
rng = fba.Rng.philox(3)
b.initiate_sampled(sim, [0], g["is/init_counts"][0][None, :], None, rng)
rs = np.random.RandomState(5)
batch = int(os.environ.get("BATCH", 4096))
pid = rs.randint(0, n, batch).astype(np.int64)
start = rs.randint(0, sim.S, batch).astype(np.int32)
depth = np.full(batch, 20, np.int32)
for coop in (0, 1):
    ctx.set_option("rollout_coop", coop)
    for _ in range(3):
        fba.rollouts(b, pid, start, depth, 0.95, rng)
    ctx.profile_begin()
    for _ in range(10):
        ret = fba.rollouts(b, pid, start, depth, 0.95, rng)
    ctx.profile_end()
    ms, k = ctx.kernel_time("k_rollouts")
    print("coop", coop, "batch", batch, "kernel ms", round(ms / k, 4), "mean ret", ret.mean())
