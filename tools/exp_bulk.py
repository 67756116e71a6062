"""Scratch experiment: full-copy gather, LDG/STG warps vs TMA bulk copies (not part of the product)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fba_pomdp_b200 as fba
import golden_util as G
g = G.load("sysadmin")
script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
n = 1_250_000
ctx = fba.Context(0)
ctx.set_option("inplace_resample", 0)
sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
for bulk in (0, 1):
    ctx.set_option("bulk_copy", bulk)
    b = fba.BAImportanceSampling(n)
    rng = fba.Rng.philox(42)
    b.initiate_sampled(sim, [0], g["is/init_counts"][0][None, :], None, rng)
    for t in range(4):
        b.updateEstimation(*script[t % len(script)], rng, want_likelihood=False)
    ctx.synchronize()
    ctx.profile_begin()
    for t in range(10):
        b.updateEstimation(*script[(4 + t) % len(script)], rng, want_likelihood=False)
    ctx.profile_end()
    kt = ctx.kernel_times()
    name = "k_gather_bulk" if bulk else "k_gather"
    ms = kt[name][0] / kt[name][1]
    # sanity: counts conserved
    d = b.download(0, 4)
    print("bulk", bulk, name, "ms", round(ms, 4), "GB/s", round(23704 * n / ms / 1e6, 1), "sum check",
          d["counts"].astype(np.float64).sum(1) - float(g["is/init_counts"][0].astype(np.float64).sum()))
    b.free()
