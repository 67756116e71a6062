#!/usr/bin/env python
"""Many runs over many GPUs: independent runs need no communication, so every rank drives its own
share of the runs (fba_b200::runBatchedExperiment on its own GPU, its own seed and its own copy of the
reference's environment); the job's throughput is the sum. Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        tools/bench_experiment_multi.py
Rank 0 prints one JSON line (time = max over ranks, between two barriers)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, sims, runs_per_gpu, episodes, horizon = 1024, 1024, 2048, 3, 20
r = pyref.Ref("episodic-tiger", horizon=horizon, seed=str(100 + rank))
r.batched_episodes(64, 4, 16, 1, device=local)  # warm-up: CUDA context, module load
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
ret, _ = r.batched_episodes(n, runs_per_gpu, sims, episodes, device=local, seed=4711 + 1000 * rank)
if world > 1:
    dist.barrier()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
stat = torch.tensor([float(ret.sum()), float(ret.size)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dist.all_reduce(stat)
r.close()
if rank == 0:
    total = runs_per_gpu * world * episodes
    print(json.dumps({"workload": "episodic tiger, %d particles, %d simulations, horizon %d, %d episodes per run, "
                                  "%d runs per GPU" % (n, sims, horizon, episodes, runs_per_gpu),
                      "n_gpus": world, "seconds": float(dt.item()), "run_episodes_per_s": total / float(dt.item()),
                      "mean_return": float(stat[0].item() / stat[1].item()),
                      "reference_one_core_run_episodes_per_s": 118.0}))
if world > 1:
    dist.destroy_process_group()
