#!/usr/bin/env python
"""Decision quality against wave width: mean episode return (reference's own episode loop and
environment, oracle/ref_harness.cpp:ref_adapter_episodes) of the reference's RBAPOUCT and of
fba_b200::CudaTreePOUCT at several wave widths, equal simulation budget. One JSON document."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref  # noqa: E402

SIMS, EPISODES, HORIZON = 1024, 300, 10
out = {"simulations": SIMS, "episodes": EPISODES, "horizon": HORIZON, "domains": {}}
for domain, kw, n in (("episodic-tiger", dict(), 512),
                      ("centered-collision-avoidance", dict(size=1, width=3, height=3, factored=True), 256)):
    r = pyref.Ref(domain, horizon=HORIZON, seed="21", **kw)
    rows = {}

    def stat(x):
        return dict(mean=float(x.mean()), se=float(x.std(ddof=1) / np.sqrt(len(x))))
    rows["random policy"] = stat(r.adapter_episodes(0, n, "random", SIMS, EPISODES))
    rows["reference RBAPOUCT (sequential)"] = stat(r.adapter_episodes(0, n, "po-uct", SIMS, EPISODES))
    for wave in (1, 16, 64, 256, 1024):
        rows["CudaTreePOUCT wave=%d" % wave] = stat(r.adapter_episodes(1, n, "cuda-tree-po-uct:%d" % wave, SIMS, EPISODES))
    r.close()
    out["domains"][domain] = rows
print(json.dumps(out, indent=1))
