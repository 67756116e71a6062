"""Scratch experiment: where does a reinvigoration update spend its time? (not part of the product)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fba_pomdp_b200 as fba
import golden_util as G
ctx = fba.Context(0)
g = G.load("ftiger_mu")
sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par, max_structures=len(g.t_par) + 20000)
n = 100_000
stride = int(g["reinv/stride"])
rs = np.random.RandomState(0)
def blow_up(prefix):
    sid, c = g[prefix + "_struct_id"], g[prefix + "_counts"]
    idx = rs.randint(0, len(sid), n)
    return dict(struct_id=sid[idx], counts=c[idx], state=rs.randint(0, sim.S, n).astype(np.int32))
b = fba.ReinvigoratingRejectionSampling(n, 1000, G.MUTATE_KIND["ftiger_mu"])
b.initiate(sim, stride=stride, belief=blow_up("reinv/init_b"), fully_connected=blow_up("reinv/init_fc"))
rng = fba.Rng.philox(2)
for t in range(3):
    b.updateEstimation(2, t % 2, rng)
ctx.synchronize()
ctx.profile_begin()
t0 = time.perf_counter()
tb = 0.0
for t in range(10):
    t1 = time.perf_counter(); b.reinvigorateParticles(rng); ctx.synchronize(); tb += time.perf_counter() - t1
    import ctypes as C
    nn = C.c_int64(0)
    for h in (b.h, b.fc):
        t2 = time.perf_counter()
        rc = b.L.fba_belief_reject_sample(h, 2, t % 2, C.byref(rng), C.byref(nn))
        print("   reject rc", rc, "attempts", nn.value, "ms", round((time.perf_counter() - t2) * 1e3, 3))
ctx.synchronize()
wall = (time.perf_counter() - t0) / 10
ctx.profile_end()
print("wall ms/update", wall * 1e3, "of which reinvigorate", tb / 10 * 1e3)
for k, v in sorted(ctx.kernel_times().items(), key=lambda kv: -kv[1][0]):
    print("  %-30s %8.3f ms/update  (%d launches)" % (k, v[0] / 10, v[1]))
