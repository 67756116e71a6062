import os, sys
import numpy as np
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fba_pomdp_b200 as fba
import golden_util as G
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_cuda_tree as T
ctx = fba.Context(0)
n = 1024
states = np.ones(n)
for kw in (dict(), dict(weighted=False), dict(delta=64)):
    g, sim, b = T._tiger(ctx, n, states, **kw)
    tree = fba.SearchTree(sim, 4096, 6)
    for rep in range(3):
        qs, vs = [], []
        for seed in range(64):
            a, q, visits = tree.selectAction(b, 4096, 3, 30.0, 0.95, 512, fba.Rng.philox(1000 * rep + 7 + seed))
            qs.append(q[2]); vs.append(visits[2])
        qs, vs = np.array(qs), np.array(vs)
        print(kw, "rep", rep, "per-seed Q(listen): mean %.2f sd %.2f min %.1f max %.1f; visits mean %.0f; weighted mean %.2f; se(mean) %.2f"
              % (qs.mean(), qs.std(ddof=1), qs.min(), qs.max(), vs.mean(), (qs * vs).sum() / vs.sum(), qs.std(ddof=1) / 8))
    tree.free(); b.free(); sim.close()
