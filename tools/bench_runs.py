"""Many independent runs batched on one GPU (fba_runs_*): belief updates per second
of R concurrent episodic-tiger runs of 1024 particles (BASELINE.json configs[0]) against (a) the same
R beliefs updated one after the other through the single-belief path and (b) the reference's own CPU
time per step at that size (update 0.36 ms + resample 1.21 ms). One JSON line.

    python tools/bench_runs.py [--particles 1024] [--steps 50] [--name tiger]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fba_pomdp_b200 as fba  # noqa: E402
import golden_util as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", default="tiger")
    ap.add_argument("--particles", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--runs", type=int, nargs="*", default=[1, 16, 148, 1024, 4096, 16384])
    ap.add_argument("--plan-runs", type=int, nargs="*", default=[1, 148, 1024, 4096])
    ap.add_argument("--sims", type=int, default=1024)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--sims-per-wave", type=int, nargs="*", default=[1, 4])
    args = ap.parse_args()
    g = G.load(args.name)
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
    ctx = fba.Context(0)
    used = np.unique(g["is/init_struct_id"])
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par[used], g.o_par[used])
    proto = g["is/init_counts"][:1]
    n = args.particles
    out = dict(workload="%s, %d particles per run, importance-sampling updateEstimation" % (args.name, n), rows=[])

    # single-belief path, one run after the other
    b = fba.BAImportanceSampling(n)
    rng = fba.Rng.philox(1)
    b.initiate_sampled(sim, [0], proto, None, rng)
    for t in range(10):
        b.updateEstimation(*script[t % len(script)], rng, want_likelihood=False)
    ctx.synchronize()
    t0 = time.perf_counter()
    for t in range(200):
        b.updateEstimation(*script[t % len(script)], rng, want_likelihood=False)
    ctx.synchronize()
    single_us = (time.perf_counter() - t0) / 200 * 1e6
    b.free()
    out["single_belief_us_per_update"] = single_us

    for R in args.runs:
        batch = fba.BatchedBAImportanceSampling(R, n)
        rng = fba.Rng.philox(2)
        batch.initiate_sampled(sim, [0], proto, None, rng)
        rs = np.random.RandomState(0)
        picks = rs.randint(0, len(script), (args.steps + 5, R))
        acts = np.array([[script[k][0] for k in row] for row in picks], np.int32)
        obs = np.array([[script[k][1] for k in row] for row in picks], np.int32)
        for t in range(5):
            batch.updateEstimation(acts[t], obs[t], rng, want_likelihood=False)
        ctx.synchronize()
        t0 = time.perf_counter()
        for t in range(5, 5 + args.steps):
            batch.updateEstimation(acts[t], obs[t], rng, want_likelihood=False)
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        # with the likelihoods read back every step (what an experiment loop does)
        t0 = time.perf_counter()
        for t in range(5, 5 + args.steps):
            batch.updateEstimation(acts[t], obs[t], rng)
        ctx.synchronize()
        dt_e2e = (time.perf_counter() - t0) / args.steps
        out["rows"].append(dict(runs=R, ms_per_batched_update=dt * 1e3, run_updates_per_s=R / dt,
                                particle_updates_per_s=R * n / dt, ms_per_update_with_likelihoods=dt_e2e * 1e3,
                                speedup_vs_one_after_the_other=single_us * 1e-6 * R / dt))
        batch.free()
    out["reference_cpu_ms_per_run_update"] = 1.57 if (args.name == "tiger" and n == 1024) else None

    # planning: Planner::selectAction of every run at once, one device tree per run (fba_runs_plan)
    out["planning"] = dict(simulations=args.sims, depth=args.depth, rows=[])
    for R in args.plan_runs:
        batch = fba.BatchedBAImportanceSampling(R, n)
        rng = fba.Rng.philox(3)
        batch.initiate_sampled(sim, [0], proto, None, rng)
        for w in args.sims_per_wave:
            batch.selectAction(min(args.sims, 64), args.depth, 5.0, 0.95, rng, sims_per_wave=w)   # warm-up, allocations
            ctx.synchronize()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                act, q, visits = batch.selectAction(args.sims, args.depth, 5.0, 0.95, rng, sims_per_wave=w)
            dt = (time.perf_counter() - t0) / reps
            out["planning"]["rows"].append(dict(runs=R, sims_per_wave=w, s_per_batched_selectAction=dt,
                                                ms_per_run_selectAction=dt / R * 1e3,
                                                simulations_per_s=R * args.sims / dt))
        batch.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
