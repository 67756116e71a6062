"""Scratch driver for ncu: the kernels of the structure-learning beliefs on sysadmin-10 (S = 1024, 20 actions) —
Gibbs state histories by message passing for 64 models over a 40-step history, the nested update for 128 top
particles with 4096 bottom states each. Prints wall times of the calls (not under ncu: plain run)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fba_pomdp_b200 as fba
import golden_util as G
from fba_pomdp_b200.structure_beliefs import NestedBelief

g = G.load("sysadmin")
ctx = fba.Context(0)
sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
S = int(g.desc["S"])
upd = [t for t in range(len(g.a)) if not (g.flags[t] & 1)]
acts = np.array([g.a[upd[k % len(upd)]] for k in range(40)], np.int32)
obs = np.array([g.o[upd[k % len(upd)]] for k in range(40)], np.int32)
hist = (np.array([20, 20], np.int32), acts, obs)
prior = np.zeros(S, np.float32)
prior[S - 1] = 1.0

n = 64
b = fba.BAImportanceSampling(n)
b.initiate(sim, struct_id=np.zeros(n, np.int32), counts=np.repeat(g["is/init_counts"][:1], n, 0), state=np.zeros(n, np.int32))
for rep in range(3):
    t0 = time.perf_counter()
    seq = b.sample_state_history("msg", *hist, fba.Rng.philox(5 + rep), state_prior=prior)
    dt = time.perf_counter() - t0
print("state histories by messages: %d models x %d steps, S = %d: %.2f ms per call (%.1f us per model and step)"
      % (n, 40, S, 1e3 * dt, 1e6 * dt / 40))
t0 = time.perf_counter()
b.add_history_counts(*hist, seq)
print("posterior counts: %.3f ms" % (1e3 * (time.perf_counter() - t0)))
b.free()

n_top, n_bot = 128, 4096
nb = NestedBelief(n_top, n_bot)
nb.initiate(sim, struct_id=np.zeros(n_top, np.int32), counts=np.repeat(g["is/init_counts"][:1], n_top, 0),
            states=np.full((n_top, n_bot), S - 1, np.int32))
for exact in (0, 1):
    ctx.set_option("nested_exact", exact)
    t0 = time.perf_counter()
    nb.updateEstimation(int(acts[0]), int(obs[0]), fba.Rng.philox(9 + exact))
    print("nested update (%s): %d x %d, attempts per top particle %.0f: %.2f ms"
          % ("thread per top particle" if exact else "warp per top particle", n_top, n_bot, nb.attempts.mean(),
             1e3 * (time.perf_counter() - t0)))
nb.free()
sim.close()
ctx.close()
