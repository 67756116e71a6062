#!/usr/bin/env python
"""selectAction latency: the reference's RBAPOUCT (CPU) vs fba_b200::CudaBatchedPOUCT (wave-parallel
POMCP, tree on the host, simulator calls batched on the GPU) vs fba_b200::CudaTreePOUCT (tree on the
device, whole simulations in one kernel), all inside the reference's own Planner interface
(oracle/ref_harness.cpp:ref_plan_seconds). BASELINE.json configs[2]: gridworld, 4096 simulations."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref

out = {}
for domain, kw, n in (("gridworld", dict(size=3), 4096), ("episodic-tiger", dict(), 4096),
                      ("linear-sysadmin", dict(size=10, factored=True), 1024)):
    r = pyref.Ref(domain, horizon=20, seed="5", **kw)
    row = {"particles": n, "simulations": 4096, "horizon": 20}
    row["reference RBAPOUCT + BAImportanceSampling (1 core), s"] = r.plan_seconds(0, n, "po-uct", 4096, 3)
    for wave in (64, 256, 1024):
        row["CudaBatchedPOUCT wave=%d + CudaBAImportanceSampling, s" % wave] = r.plan_seconds(
            1, n, "cuda-po-uct:%d" % wave, 4096, 3)
    for wave in (64, 256, 1024, 4096):
        row["CudaTreePOUCT (tree on device) wave=%d + CudaBAImportanceSampling, s" % wave] = r.plan_seconds(
            1, n, "cuda-tree-po-uct:%d" % wave, 4096, 5)
    r.close()
    out[domain] = row
print(json.dumps(out, indent=1))
