#!/usr/bin/env python
"""Whole-experiment throughput: fba_b200::runBatchedExperiment (R runs of BA-POMCP in lockstep on one
GPU: beliefs in one fba_runs object, one device search tree per run, sequential search per run)
against the reference's own experiment loop (BAImportanceSampling + RBAPOUCT, one run after the other
on one core). Same domain, particles, simulations, horizon, episodes. One JSON document."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref  # noqa: E402

out = {}
for domain, kw, n, sims, runs, ref_runs in (("episodic-tiger", dict(), 1024, 1024, 2048, 100),
                                           ("gridworld", dict(size=3), 256, 512, 1024, 100)):
    horizon, episodes = 20, 3
    r = pyref.Ref(domain, horizon=horizon, seed="9", **kw)
    r.batched_episodes(64, 4, 16, 1)  # warm-up: CUDA context, module load
    ours, dt = r.batched_episodes(n, runs, sims, episodes)
    t0 = time.perf_counter()
    ref = np.stack([r.adapter_episodes(0, n, "po-uct", sims, episodes) for _ in range(ref_runs)], axis=1)
    dt_ref = time.perf_counter() - t0
    r.close()
    out[domain] = {
        "particles": n, "simulations": sims, "horizon": horizon, "episodes": episodes,
        "batched": {"runs": runs, "seconds": dt, "run_episodes_per_s": runs * episodes / dt,
                    "mean_return_per_episode": ours.mean(1).tolist(),
                    "se": (ours.std(1, ddof=1) / np.sqrt(runs)).tolist()},
        "reference (1 core)": {"runs": ref_runs, "seconds": dt_ref, "run_episodes_per_s": ref_runs * episodes / dt_ref,
                               "mean_return_per_episode": ref.mean(1).tolist(),
                               "se": (ref.std(1, ddof=1) / np.sqrt(ref_runs)).tolist()},
    }
    out[domain]["speedup_vs_one_core"] = out[domain]["batched"]["run_episodes_per_s"] / out[domain]["reference (1 core)"]["run_episodes_per_s"]
print(json.dumps(out, indent=1))
