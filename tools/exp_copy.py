"""Scratch experiment: why is k_copy_inplace slower on the sharded path? (not part of the product)"""
import ctypes as C
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import fba_pomdp_b200 as fba
import golden_util as G

g = G.load("sysadmin")
script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
n = 1_250_000
ctx = fba.Context(0)
sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
proto = g["is/init_counts"][0]

def run(kind):
    b = fba.BAImportanceSampling(n)
    rng = fba.Rng.philox(42)
    b.initiate_sampled(sim, [0], proto[None, :], None, rng)
    L, h = b.L, b.h
    def step(t):
        a, o = script[t % len(script)]
        if kind == "single":
            b.updateEstimation(a, o, rng, want_likelihood=False)
        elif kind == "phases-host-total":
            loc = C.c_double(0)
            assert L.fba_belief_propose(h, a, o, C.byref(rng), C.byref(loc)) == 0
            tot = np.array([loc.value])
            assert L.fba_belief_shard_resample(h, fba.capi.ptr(tot), 1, 0, 0.37, C.byref(rng), None, None) == 0
        elif kind == "phases-normalize-only":
            loc = C.c_double(0)
            assert L.fba_belief_propose(h, a, o, C.byref(rng), C.byref(loc)) == 0
            assert L.fba_belief_normalize(h, loc.value) == 0
            assert L.fba_belief_resample(h, C.byref(rng)) == 0
    for t in range(5):
        step(t)
    ctx.synchronize()
    ctx.profile_begin()
    for t in range(10):
        step(5 + t)
    ctx.profile_end()
    kt = ctx.kernel_times()
    print(kind, {k: round(v[0] / v[1], 4) for k, v in kt.items() if k.startswith(("k_copy", "k_offspring", "k_scale"))},
          "copies", b.resample_stats())
    b.free()

for kind in sys.argv[1:] or ["single", "phases-host-total", "phases-normalize-only", "single"]:
    run(kind)
