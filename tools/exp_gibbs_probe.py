import os, sys, time
sys.path.insert(0, "/root/repo/oracle")
import pyref
CA = dict(size=1, width=5, height=5, factored=True, structure_prior="match-uniform")
FT = dict(size=4, factored=True, structure_prior="match-uniform")
for name, dom, kw, kind, n, ep in (("gibbs-msg CA", "centered-collision-avoidance", CA, 15, 64, 6),
                                   ("gibbs-msg FT", "episodic-factored-tiger", FT, 15, 256, 30),
                                   ("cheat CA", "centered-collision-avoidance", CA, 9, 65536, 10)):
    for rep in range(3):
        r = pyref.Ref(dom, horizon=12, seed="5", **kw)
        r.adapter_episodes(kind, n, "random", 1, 1)
        t0 = time.perf_counter()
        r.adapter_episodes(kind, n, "random", 1, ep)
        wall = time.perf_counter() - t0
        s, calls = r.adapter_update_seconds()
        print(name, "rep", rep, "updates", calls, "ms/update %.3f" % (1e3 * s / max(calls, 1)), "events", r.adapter_events(), "wall %.2f s" % wall, flush=True)
        r.close()
