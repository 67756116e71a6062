"""Scratch experiment: k_propose time vs occupancy / L2 fetch granularity (not part of the product)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fba_pomdp_b200 as fba
import golden_util as G
g = G.load("sysadmin")
script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
n = 1_250_000
ctx = fba.Context(0)
sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
b = fba.BAImportanceSampling(n)
rng = fba.Rng.philox(42)
b.initiate_sampled(sim, [0], g["is/init_counts"][0][None, :], None, rng)
for t in range(5):
    b.updateEstimation(*script[t % len(script)], rng, want_likelihood=False)
ctx.synchronize()
ctx.profile_begin()
for t in range(12):
    b.updateEstimation(*script[(5 + t) % len(script)], rng, want_likelihood=False)
ctx.profile_end()
kt = ctx.kernel_times()
print(os.environ.get("FBA_B200_LIB", "default").split("/")[-1], 
      {k: round(v[0] / v[1], 4) for k, v in kt.items() if k.startswith(("k_copy", "k_propose"))})
