#!/usr/bin/env python
"""Throughput of the CUDA path on all five BASELINE.json configurations (one GPU, PHILOX mode),
with the reference's CPU numbers from BASELINE.md §2 beside them. Writes one JSON object.
Not the driver's bench (that is bench.py, which runs configs[4]); this is the per-config table
DESIGN.md §6 quotes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fba_pomdp_b200 as fba
import golden_util as G


def prototypes(g, prefix="is/init"):
    """Distinct (structure id, count block) pairs the reference prior produced in the fixture."""
    sid, counts = g[prefix + "_struct_id"], g[prefix + "_counts"]
    seen, psid, pc = {}, [], []
    for i in range(len(sid)):
        k = (int(sid[i]), counts[i].tobytes())
        if k not in seen:
            seen[k] = len(psid)
            psid.append(int(sid[i]))
            pc.append(counts[i])
    return np.array(psid, np.int32), np.stack(pc)


def script_of(g):
    return [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]


def time_updates(ctx, fn, steps, warmup=3):
    for t in range(warmup):
        fn(t)
    ctx.synchronize()
    t0 = time.perf_counter()
    for t in range(steps):
        fn(warmup + t)
    ctx.synchronize()
    return (time.perf_counter() - t0) / steps


def is_config(ctx, name, n, steps=20, delta_capacity=0):
    g = G.load(name)
    sim = fba.BAPOMDP(ctx, dict(g.desc, delta_capacity=delta_capacity), g.t_par, g.o_par)
    psid, pc = prototypes(g)
    b = fba.BAImportanceSampling(n)
    rng = fba.Rng.philox(1)
    b.initiate_sampled(sim, psid, pc, np.ones(len(psid)), rng, stride=pc.shape[1])
    sc = script_of(g)
    dt = time_updates(ctx, lambda t: b.updateEstimation(*sc[t % len(sc)], rng, want_likelihood=False), steps)
    copies, res = b.resample_stats()
    out = {"particles": n, "structures": int(len(psid)), "cells_per_particle": int(pc.shape[1]),
           "ms_per_update": dt * 1e3, "particles_per_s": n / dt,
           "copied_fraction": copies / max(res, 1) / n}
    b.free()
    sim.close()
    return out


def main():
    ctx = fba.Context(0)
    out = {}
    # configs[0]: episodic tiger, tabular, IS, 1024 particles (latency-bound: ~10 launches)
    out["1 episodic-tiger BA-POMDP IS N=1024"] = dict(
        is_config(ctx, "tiger", 1024, 200),
        reference_cpu="update 0.36 ms + resample 1.21 ms per step = 6.5e5 particles/s (BASELINE.md §2)")
    # configs[1]: factored tiger size 8, 1e5 particles per filter + reinvigoration (resample 1000)
    g = G.load("ftiger_mu")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par, max_structures=len(g.t_par) + 20000)
    n = 100_000
    stride = int(g["reinv/stride"])
    rs = np.random.RandomState(0)

    def blow_up(prefix):
        sid, st, c = g[prefix + "_struct_id"], g[prefix + "_state"], g[prefix + "_counts"]
        idx = rs.randint(0, len(sid), n)
        return dict(struct_id=sid[idx], counts=c[idx], state=rs.randint(0, sim.S, n).astype(np.int32))

    b = fba.ReinvigoratingRejectionSampling(n, 1000, G.MUTATE_KIND["ftiger_mu"])
    b.initiate(sim, stride=stride, belief=blow_up("reinv/init_b"), fully_connected=blow_up("reinv/init_fc"))
    rng = fba.Rng.philox(2)
    dt = time_updates(ctx, lambda t: b.updateEstimation(2, t % 2, rng), 10, 2)  # listen, alternating growls
    out["2 factored-tiger-8 FBA-POMDP reinvigoration N=1e5 (+1e5 fully connected), resample 1000"] = {
        "particles": n, "ms_per_update": dt * 1e3, "particles_per_s": 2 * n / dt,
        "structures_after": int(sim.num_structures),
        "reference_cpu": "IS at N=4096: 3.0e5 (update) / 7.6e4 (resample) particles/s (BASELINE.md §2)"}
    b.free()
    sim.close()
    out["2b factored-tiger-8 FBA-POMDP IS N=1e5 (match-uniform structures)"] = is_config(ctx, "ftiger_mu", 100_000)
    # configs[2]: gridworld size 3, 1e6 particles (rollouts are in bench.py's rollouts leg)
    out["3 gridworld-3 BA-POMDP IS N=1e6"] = dict(
        is_config(ctx, "gridworld3", 1_000_000, 10),
        reference_cpu="size 5 at N=4096: 2.5e5 / 3.9e4 particles/s (BASELINE.md §2)")
    # the same at --size 5 (720 KB dense per particle): base+delta storage, 256 increments per particle
    out["3b gridworld-5 BA-POMDP IS N=1e6, base+delta storage"] = dict(
        is_config(ctx, "gridworld5", 1_000_000, 10, delta_capacity=256),
        reference_cpu="N=4096: 2.5e5 / 3.9e4 particles/s (BASELINE.md §2)")
    # configs[3]: collision avoidance 5x5x1, heterogeneous structures, 1e6 particles
    out["4 collision-avoidance 5x5x1 FBA-POMDP IS N=1e6 (match-uniform structures)"] = dict(
        is_config(ctx, "ca", 1_000_000),
        reference_cpu="N=4096: 5.9e5 / 8.9e4 particles/s (BASELINE.md §2)")
    # configs[4]: sysadmin, one GPU's shard of 1e7
    out["5 linear-sysadmin-10 FBA-POMDP IS N=1.25e6 (1/8 of 1e7)"] = dict(
        is_config(ctx, "sysadmin", 1_250_000),
        reference_cpu="N=4096: 2.2e5 / 1.1e4 particles/s (BASELINE.md §2)")
    ctx.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
