#!/usr/bin/env python
"""Belief-update cost of the structure-learning beliefs: the reference's own classes (CPU, one
core) against this repo's CUDA adapters, both inside the reference's own episode loop (oracle/ref_harness.cpp:
ref_adapter_episodes, `random` planner), the wall time measured INSIDE Belief::updateEstimation by a forwarding
wrapper, call by call. For the MH beliefs the mean includes the MH / Gibbs runs the log-likelihood threshold (-4)
triggers; `events` counts them on the CUDA side. The first calls of a fresh belief (first-use allocations) are
reported separately."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref

HORIZON = 12
FT = dict(size=4, factored=True, structure_prior="match-uniform")
CA = dict(size=1, width=5, height=5, factored=True, structure_prior="match-uniform")
CASES = [
    # name, domain, kwargs, (reference kind, cuda kind), particles reference / CUDA, episodes reference / CUDA
    ("MHNIPS2018", "episodic-factored-tiger", FT, (6, 7), 256, 256, 30, 30),
    ("MHNIPS2018", "episodic-factored-tiger", FT, (None, 7), None, 4096, 0, 30),
    ("MHwithinGibbs (messages)", "episodic-factored-tiger", FT, (14, 15), 256, 256, 30, 30),
    ("MHwithinGibbs (rejection)", "episodic-factored-tiger", FT, (16, 17), 256, 256, 30, 30),
    ("MHwithinGibbs (messages)", "centered-collision-avoidance", CA, (14, 15), 64, 64, 12, 12),
    ("CheatingReinvigoration", "episodic-factored-tiger", FT, (8, 9), 1024, 65536, 20, 20),
    ("CheatingReinvigoration", "centered-collision-avoidance", CA, (8, 9), 512, 65536, 20, 20),
    ("StructureIncubatorSampling", "episodic-factored-tiger", FT, (10, 11), 1024, 65536, 20, 20),
    ("StructureIncubatorSampling", "linear-sysadmin", dict(size=6, factored=True), (10, 11), 256, 16384, 2, 10),
    ("NestedBelief (n, n^2)", "episodic-tiger", dict(), (12, 13), 24, 24, 20, 20),
    ("NestedBelief (n, n^2)", "episodic-tiger", dict(), (None, 13), None, 128, 0, 20),
    ("NestedBelief (n, n^2)", "episodic-factored-tiger", dict(size=3, factored=True, structure_prior="match-uniform"),
     (12, 13), 16, 16, 20, 20),
]

for name, domain, kw, (k_ref, k_cuda), n_ref, n_cuda, e_ref, e_cuda in CASES:
    r = pyref.Ref(domain, horizon=HORIZON, seed="5", **kw)
    row = {"belief": name, "domain": domain}
    try:
        for tag, kind, n, episodes in (("reference", k_ref, n_ref, e_ref), ("cuda", k_cuda, n_cuda, e_cuda)):
            if kind is None:
                continue
            if tag == "cuda":
                r.adapter_episodes(kind, n, "random", 1, 1)           # warm-up: context, first-use allocations
            r.adapter_episodes(kind, n, "random", 1, episodes)
            t = r.adapter_update_times()
            # a fresh belief pays its first-use allocations (back buffer, rejection-sampling waves, scratch) in its
            # first calls: the steady state is the mean over the calls after the first SKIP — MH / Gibbs runs and
            # cheats, which are rare and expensive, stay in it — next to the median (a plain update)
            SKIP = 4
            steady = t[SKIP:] if len(t) > 2 * SKIP else t
            row[tag] = {"particles": n, "updates": int(len(t)), "ms_per_update": 1e3 * float(steady.mean()),
                        "median_ms_per_update": 1e3 * float(np.median(steady)),
                        "first_calls_ms": [round(1e3 * float(x), 3) for x in t[:SKIP]],
                        "particle_updates_per_s": n / float(steady.mean())}
            if tag == "cuda":
                row[tag]["events"] = r.adapter_events()
        if "reference" in row:
            row["speedup_per_particle_update"] = (row["cuda"]["particle_updates_per_s"]
                                                  / row["reference"]["particle_updates_per_s"])
    finally:
        r.close()
    print(json.dumps(row), flush=True)
