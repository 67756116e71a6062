#!/usr/bin/env python
"""bench.py — particle belief updates/s on B200 for BASELINE.json's headline configuration.

Workload (config.workload): linear-sysadmin, 10 computers, FBA-POMDP (configs[4]); the belief is
importance sampling over N particles, one "step" = one full belief update of every particle
(importance_sampling::update + ::resample) on a fixed (action, observation) script taken from the
reference's true environment (tests/golden/sysadmin.npz). BASELINE.json shards 10^7 particles over
8 GPUs; weak scaling keeps that per-GPU shard (1.25e6 particles, 14.8 GB of counts, far above the
126 MB L2) fixed, so N=1 runs one shard.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...             # the reference's own CPU path
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "particle belief updates/s"
UNIT = "particles/s"
WORKLOAD = "linear-sysadmin-10 FBA-POMDP, importance-sampling belief update (update+resample)"
PER_GPU_PARTICLES = 1_250_000  # 10^7 / 8 (BASELINE.json configs[4])


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


_POLLER = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
print("max %f" % nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)
out = []
import select
while True:
    out.append((time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
    if select.select([sys.stdin], [], [], 0.002)[0]:
        break
for t, mhz, mask in out:
    print("%.6f %d %d" % (t, mhz, mask))
"""


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region, sampled by a SEPARATE PROCESS (NVML
    polled every 2 ms — the counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*`
    prints; nvidia-smi itself cannot start and sample inside a ~25 ms region). The poller is started
    well before the region and keeps (wall time, MHz, reasons) for every sample; the bench process
    only notes the wall-clock window of the region and filters afterwards, so no Python thread of the
    timed process competes with the launches."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.proc, self.max_mhz, self.windows = None, None, []
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _POLLER, str(phys)], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            first = self.proc.stdout.readline().split()
            self.max_mhz = float(first[1]) if len(first) == 2 and first[0] == "max" else None
            if self.max_mhz is None:
                self.proc = None
        except Exception:  # noqa: BLE001
            self.proc = None

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def finish(self):
        """-> one clocks dict per window"""
        rows = []
        if self.proc is not None:
            try:
                out, _ = self.proc.communicate("stop\n", timeout=20)
                rows = [tuple(float(x) for x in ln.split()) for ln in out.splitlines() if len(ln.split()) == 3]
            except Exception:  # noqa: BLE001
                self.proc.kill()
        res = []
        for t0, t1 in self.windows:
            sel = [r for r in rows if t0 <= r[0] <= t1]
            if not sel:  # region shorter than the polling period: nearest samples either side
                sel = sorted(rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:2]
            reasons = sorted(n for n, bit in self.REASONS.items() if any(int(r[2]) & bit for r in sel))
            res.append({"sm_mhz": float(np.median([r[1] for r in sel])) if sel else None,
                        "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sel),
                        "sampler": "separate process, NVML every 2 ms"})
        return res


WORKLOADS = {
    # BASELINE.json configs[4] (the headline): one shared structure
    "sysadmin": dict(fixture="sysadmin", particles=PER_GPU_PARTICLES, text=WORKLOAD,
                     ref=dict(domain="linear-sysadmin", size=10, factored=True)),
    # configs[3]: collision avoidance 5x5, 1 obstacle, --structure-prior match-uniform
    # (heterogeneous DBN structures), 10^6 particles per GPU
    "ca": dict(fixture="ca", particles=1_000_000,
               text="centered-collision-avoidance 5x5x1 FBA-POMDP (match-uniform structure prior), "
                    "importance-sampling belief update (update+resample)",
               ref=dict(domain="centered-collision-avoidance", size=1, width=5, height=5, factored=True,
                        structure_prior="match-uniform")),
}


def workload(name="sysadmin"):
    """(fixture, prototype structure ids, prototype count blocks, (a,o) script): the distinct
    (structure, counts) pairs the reference's prior produced become equally likely prototypes."""
    import golden_util as G
    g = G.load(WORKLOADS[name]["fixture"])
    sid, counts = g["is/init_struct_id"], g["is/init_counts"]
    seen, psid, pc = {}, [], []
    for i in range(len(sid)):
        k = (int(sid[i]), counts[i].tobytes())
        if k not in seen:
            seen[k] = len(psid)
            psid.append(int(sid[i]))
            pc.append(counts[i])
    steps = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
    return g, np.array(psid, np.int32), np.stack(pc), steps


# ------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------
def _ref_worker(n, steps, warmup, seed, script, q, wl="sysadmin"):
    try:
        import pyref
        kind = "reference"
        kw = dict(WORKLOADS[wl]["ref"])
        r = pyref.Ref(kw.pop("domain"), seed=seed, **kw)
        r.belief_init(pyref.F_IS, n)

        def one(t):
            a, o = script[t % len(script)]
            r.update_estimation(pyref.F_IS, a, o)
    except Exception:  # the compiled reference did not travel: time the C port of it instead
        kind = "port"
        import golden_util as G
        import pyoracle as O
        g = G.load(WORKLOADS[wl]["fixture"])
        m = O.Model(g.desc)
        st = O.Structs(m, g.t_par, g.o_par)
        state = {"b": O.Belief(n, g["is/init_counts"].shape[1])}
        state["b"].counts[:] = g["is/init_counts"][0]
        state["b"].state[:] = g["is/init_state"][0]
        state["b"].total_weight = O.sequential_uniform_total(n)
        words = np.random.RandomState(int(seed)).randint(0, 2**32, size=(2 * 11 + 2) * n + 64,
                                                         dtype=np.uint64).astype(np.uint32)

        def one(t):
            a, o = script[t % len(script)]
            rng = O.Rng(words)
            O.is_update(m, st, state["b"], a, o, rng)
            state["b"], _ = O.is_resample(state["b"], rng)

    for t in range(warmup):
        one(t)
    t0 = time.perf_counter()
    for t in range(steps):
        one(warmup + t)
    q.put((kind, time.perf_counter() - t0))


def run_reference_cpu(n_particles, steps, warmup, replicas, wl="sysadmin"):
    """`replicas` independent single-threaded reference beliefs (the reference has no threads),
    one per host core; returns (particles/s aggregate, seconds per step, kind)."""
    import multiprocessing as mp
    script = workload(wl)[3]
    ctxmp = mp.get_context("fork")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_ref_worker, args=(n_particles, steps, warmup, str(42 + i), script, q, wl))
             for i in range(replicas)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    slowest = max(t for _, t in res)
    return replicas * n_particles * steps / slowest, slowest / steps, res[0][0]


def rollouts_leg(ctx, fba, args, torch, world=1, rank=0, dist=None):
    """BASELINE.json configs[2]: gridworld BA-POMDP, 10^6 particles (sharded over the GPUs), 4096 batched
    random-policy rollouts per planning step (RBAPOUCT::rollout x 4096, root-parallel over the ranks),
    plus the saturated rate at 2^20 rollouts per launch per GPU. Host requests in, host returns out (that
    is the call a planner makes). roofline: bytes = 4 (S + O) per simulated step (one phi
    row + one psi row) x the steps the kernel actually executed (counted on the device: rollouts end
    early at terminal states) / the kernel's CUDA-event time."""
    import golden_util as G
    g = G.load("gridworld3")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    n_total = args.rollout_particles
    n = n_total // world
    peak, peak_src = load_peaks()
    shared = np.random.RandomState(11)
    if world > 1:
        b = fba.ShardedBAImportanceSampling(n, exchange=args.exchange)
        rng = b.rank_rng(args.seed + 1)
    else:
        b = fba.BAImportanceSampling(n)
        rng = fba.Rng.philox(args.seed + 1)
    b.initiate_sampled(sim, [0], g["is/init_counts"][0][None, :], None, rng)
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
    for t in range(2):  # a learned (non-prior) belief
        if world > 1:
            b.updateEstimation(script[t][0], script[t][1], rng, step_uniform=float(shared.random_sample()),
                               likelihood=False)
        else:
            b.updateEstimation(script[t][0], script[t][1], rng, want_likelihood=False)
    depth, disc = int(g.horizon), float(g.discount)
    bytes_per_step = 4 * (sim.S + sim.O)

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.synchronize()

    def one(total, gather):
        """`total` requests over all ranks: root particles drawn from the belief on the device, their
        domain states gathered, one k_rollouts launch per rank"""
        if world > 1:
            return b.rollouts(total, depth, disc, rng, gather=gather)
        idx = np.zeros(total, np.int64)
        st = np.zeros(total, np.int32)
        import ctypes as C
        from fba_pomdp_b200.beliefs import _check
        from fba_pomdp_b200.capi import ptr
        _check(ctx.h, b.L.fba_belief_sample_batch(b.h, C.byref(rng), total, ptr(idx)))
        _check(ctx.h, b.L.fba_belief_gather_states(b.h, total, ptr(idx), ptr(st)))
        return fba.rollouts(b, idx, st, np.full(total, depth, np.int32), disc, rng)

    out = {}
    for tag, total, reps, gather in (("batch_4096", 4096, 50, True), ("batch_1048576_per_gpu", (1 << 20) * world, 5, False)):
        for _ in range(3):
            ret = one(total, gather)
        barrier()
        steps0 = ctx.counter(0)
        ctx.profile_begin()
        t0 = time.perf_counter()
        for _ in range(reps):
            ret = one(total, gather)
        barrier()
        wall = time.perf_counter() - t0
        ctx.profile_end()
        kms, kn = ctx.kernel_time("k_rollouts")
        steps = ctx.counter(0) - steps0            # this rank's share
        if world > 1:
            tt = torch.tensor([wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            wall = float(tt.item())
        ach = bytes_per_step * steps / (kms * 1e-3) / 1e9
        out[tag] = {"rollouts_per_s_e2e": total * reps / wall,
                    "rollouts_per_s_kernel_per_gpu": (total / world) / (kms / kn * 1e-3),
                    "ms_per_batch_e2e": wall / reps * 1e3, "ms_per_batch_kernel": kms / kn,
                    "simulated_steps_per_rollout": steps / float(reps * total / world),
                    "mean_return": float(np.mean(ret)),
                    "roofline": {"bound": "hbm", "kernel": "k_rollouts", "achieved": ach, "peak": peak, "unit": "GB/s",
                                 "frac": ach / peak, "peak_source": peak_src,
                                 "algorithmic_bytes_per_launch": bytes_per_step * steps / float(kn),
                                 "traffic": None,
                                 "note": "4 (S + O) = %d B per simulated step x steps executed (device counter); "
                                         "latency-bound: dependent row scans of one thread per rollout" % bytes_per_step}}
    out["roofline"] = out["batch_1048576_per_gpu"]["roofline"]
    out.update(workload="gridworld --size 3 tabular BA-POMDP (S=27, A=4, O=27, 5832 count cells/particle), "
                        "%d particles over %d GPU(s), depth %d, discount %.2f; root particles drawn from the belief "
                        "on the device, requests split evenly over the ranks" % (n * world, world, depth, disc),
               n_gpus=world, unit="rollouts/s")
    # the reference's RBAPOUCT::rollout on one host core, bounded sample
    if rank == 0:
        try:
            import pyref
            r = pyref.Ref("gridworld", size=3, horizon=g.horizon, discount=g.discount, seed="42")
            r.belief_init(pyref.F_IS, 256)
            m = 4096
            rs = np.random.RandomState(5)
            pid, start = rs.randint(0, 256, m), rs.randint(0, sim.S, m)
            t0 = time.perf_counter()
            for i in range(m):
                r.rollout(pyref.F_IS, int(pid[i]), int(start[i]), g.horizon)
            out["cpu_baseline"] = {"value": m / (time.perf_counter() - t0), "unit": "rollouts/s", "cores": 1,
                                   "kind": "reference", "sample": "4096 x RBAPOUCT::rollout, depth 20, 1 thread"}
            r.close()
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"unavailable": str(e)[:200]}
    b.free()
    sim.close()
    return out


def many_runs_leg(ctx, fba, args):
    """BASELINE.json configs[0] (episodic tiger, 1024 particles) the way the reference's real workloads
    use it — thousands of independent runs — batched on one GPU (fba_runs_*): belief updates of all
    runs in one launch (one CTA per run, bit-identical to separate beliefs), and POMCP planning with one
    device tree per run, each run searching sequentially (one simulation per run per wave)."""
    import golden_util as G
    g = G.load("tiger")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    R, n, sims = 4096, 1024, 256
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)]
    batch = fba.BatchedBAImportanceSampling(R, n)
    rng = fba.Rng.philox(args.seed + 2)
    batch.initiate_sampled(sim, [0], g["is/init_counts"][:1], None, rng)
    rs = np.random.RandomState(1)
    steps = 30
    picks = rs.randint(0, len(script), (steps + 3, R))
    acts = np.array([[script[k][0] for k in row] for row in picks], np.int32)
    obs = np.array([[script[k][1] for k in row] for row in picks], np.int32)
    for t in range(3):
        batch.updateEstimation(acts[t], obs[t], rng, want_likelihood=False)
    ctx.synchronize()
    l0 = ctx.launches
    t0 = time.perf_counter()
    for t in range(3, 3 + steps):
        batch.updateEstimation(acts[t], obs[t], rng)     # host (a, o) in, likelihoods out, every step
    dt = (time.perf_counter() - t0) / steps
    launches = (ctx.launches - l0) / steps
    batch.selectAction(sims, g.horizon, 5.0, g.discount, rng)   # allocations (tree tables) + warm-up
    ctx.synchronize()
    t0 = time.perf_counter()
    act, q, visits = batch.selectAction(sims, g.horizon, 5.0, g.discount, rng, sims_per_wave=1)
    dp = time.perf_counter() - t0
    assert int(visits.sum()) == R * sims
    batch.free()
    sim.close()
    return {"workload": "episodic tiger, %d particles per run, %d runs batched on one GPU" % (n, R),
            "belief_updates": {"ms_per_batched_update_e2e": dt * 1e3, "run_updates_per_s": R / dt,
                               "particle_updates_per_s": R * n / dt, "launches_per_update": launches,
                               "reference_cpu_ms_per_run_update": 1.57},
            "planning": {"simulations_per_run": sims, "depth": int(g.horizon), "sims_per_run_per_wave": 1,
                         "s_per_batched_selectAction": dp, "simulations_per_s": R * sims / dp,
                         "reference_cpu_simulations_per_s": 3.6e5}}


def structure_beliefs_leg(ctx, fba, args):
    """The device bricks of the reference's structure-learning beliefs on sysadmin-10,
    through the public calls with host arguments and host results (wall clock, median of 5 after a warm-up):
    MHwithinGibbs' state histories by message passing and its posterior counts for 64 chains, LogBDScore, and a
    NestedBelief update (128 top particles x 4096 bottom states)."""
    import golden_util as G
    from fba_pomdp_b200.structure_beliefs import NestedBelief
    g = G.load("sysadmin")
    sim = fba.BAPOMDP(ctx, g.desc, g.t_par, g.o_par)
    S = int(g.desc["S"])
    upd = [t for t in range(len(g.a)) if not (g.flags[t] & 1)]
    acts = np.array([g.a[upd[k % len(upd)]] for k in range(40)], np.int32)
    obs = np.array([g.o[upd[k % len(upd)]] for k in range(40)], np.int32)
    hist = (np.array([20, 20], np.int32), acts, obs)
    prior = np.zeros(S, np.float32)
    prior[S - 1] = 1.0

    def timed(fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts)) * 1e3

    chains = 64
    b = fba.BAImportanceSampling(chains)
    proto = np.repeat(g["is/init_counts"][:1], chains, 0)
    b.initiate(sim, struct_id=np.zeros(chains, np.int32), counts=proto, state=np.zeros(chains, np.int32))
    pb = fba.BAImportanceSampling(chains)
    pb.initiate(sim, struct_id=np.zeros(chains, np.int32), counts=proto, state=np.zeros(chains, np.int32))
    rng = fba.Rng.philox(args.seed + 5)
    seq = [None]

    def histories():
        seq[0] = b.sample_state_history("msg", *hist, rng, state_prior=prior)
    ms_hist = timed(histories)
    ms_counts = timed(lambda: b.add_history_counts(*hist, seq[0]))
    ms_score = timed(lambda: fba.log_bd_score(b, pb))
    b.free(), pb.free()

    n_top, n_bot = 128, 4096
    nb = NestedBelief(n_top, n_bot)
    nb.initiate(sim, struct_id=np.zeros(n_top, np.int32), counts=np.repeat(g["is/init_counts"][:1], n_top, 0),
                states=np.full((n_top, n_bot), S - 1, np.int32))
    ms_nested = timed(lambda: nb.updateEstimation(int(acts[0]), int(obs[0]), rng))
    attempts = float(nb.attempts.mean())
    nb.free()
    sim.close()
    return {"workload": "linear-sysadmin-10 (S = 1024, 20 actions), history of 2 episodes x 20 steps",
            "gibbs_state_histories_by_messages": {"chains": chains, "ms_per_call": ms_hist,
                                                  "us_per_history_step_all_chains_in_parallel": 1e3 * ms_hist / 40},
            "gibbs_posterior_counts": {"chains": chains, "ms_per_call": ms_counts},
            "log_bd_score": {"models": chains, "ms_per_call": ms_score},
            "nested_update": {"top": n_top, "bottom": n_bot, "ms_per_update": ms_nested,
                              "attempts_per_top_particle": attempts,
                              "bottom_particle_attempts_per_s": n_top * attempts / (ms_nested * 1e-3)},
            "note": "parity: tests/test_cuda_gibbs.py, tests/test_cuda_composite.py (REPLAY bit-exact vs the oracle pinned "
                    "to the reference's classes); whole-belief update cost against the reference's core: "
                    "profiles/r2z_structure_beliefs.jsonl"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # the CPU arm runs once per box
    cores = os.cpu_count() or 1
    n = args.ref_particles
    value, s_per_step, kind = run_reference_cpu(n, args.steps, args.warmup, cores, args.workload)
    sample = ("%d independent single-threaded replicas (one per host core), each a BAImportanceSampling "
              "belief of %d particles (the reference's resample is O(N^2); BASELINE.md quotes it at "
              "N=4096), %d updateEstimation steps" % (cores, n, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 counts / f64 weights",
        "data": "synthetic", "config": {"workload": WORKLOADS[args.workload]["text"], "particles_per_replica": n,
                                        "replicas": cores, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# this repo's CUDA path
# ------------------------------------------------------------------------------------------------
def belief_leg(wl, args, fba, torch, dist, ctx, world, rank, sampler, steps, headline, journal_updates=0):
    """The importance-sampling belief update of workload `wl` on `world` GPUs. Three passes over the same
    belief, each bracketed by barrier + synchronize:
      1. VALUE: K updates, nothing waits for the GPU, one CUDA event per step on the launching stream
         (per-step min / median / max), device time = first to last event, max over ranks;
      2. KERNELS: the same updates with a CUDA-event pair around every launch (per-kernel table, roofline);
      3. E2E: through the public call with host arguments and host results — per step (a, o) go in as
         call arguments, the GLOBAL step likelihood comes back (8 B D2H, one stream synchronisation),
         and Belief::sample() materialises one particle on the host of the rank that owns it."""
    g, psid, protos, script = workload(wl)
    n_local = args.particles or WORKLOADS[wl]["particles"]
    J_nodes = len(np.asarray(g.desc["feat_s"]).reshape(-1)) + len(np.asarray(g.desc["feat_o"]).reshape(-1))
    journal_updates = journal_updates or args.journal_updates
    desc = dict(g.desc, delta_capacity=J_nodes * journal_updates) if journal_updates else g.desc
    sim = fba.BAPOMDP(ctx, desc, g.t_par, g.o_par)
    # count cells a particle owns (mean over the prior's structures when they differ)
    C = int(round(np.mean([sim.structure_size(int(i)) for i in psid])))
    bytes_per_particle = 2 * (4 * C + 4 + 8)  # counts + state + weight, read + written
    sharded = world > 1 or args.force_sharded
    if sharded:
        b = fba.ShardedBAImportanceSampling(n_local, exchange=args.exchange)
        rng = b.rank_rng(args.seed)
    else:
        b = fba.BAImportanceSampling(n_local)
        rng = fba.Rng.philox(args.seed)
    b.initiate_sampled(sim, psid, protos, None if len(psid) == 1 else np.ones(len(psid)), rng,
                       stride=protos.shape[1])
    shared = np.random.RandomState(args.seed)  # same on every rank: quota offsets, owner of the sampled particle

    age = [0]  # updates this belief has been through (= updates every particle's journal holds)

    def step(t, likelihood=False):
        a, o = script[t % len(script)]
        age[0] += 1
        if sharded:
            return b.updateEstimation(a, o, rng, step_uniform=float(shared.random_sample()), likelihood=likelihood)
        return b.updateEstimation(a, o, rng, want_likelihood=likelihood)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    # setup (untimed, before the W warm-up steps): first-use costs that are not part of a step
    for t in range(2):
        step(t)
    for t in range(args.warmup):
        step(t)
    barrier()

    FS, FO = sim.FS, sim.FO
    fs_sizes = [int(x) for x in np.asarray(g.desc["feat_s"]).reshape(-1)]
    fo_sizes = [int(x) for x in np.asarray(g.desc["feat_o"]).reshape(-1)]
    # k_propose per particle: each sampled row read once, one cell written per node, the
    # likelihood rows, state (r+w), weight (r+w), structure id
    bytes_propose = sum(4 * r + 4 for r in fs_sizes) + sum(2 * 4 * r + 4 for r in fo_sizes) + 8 + 16 + 4
    peak, peak_src = load_peaks()
    stream = torch.cuda.ExternalStream(ctx.stream)

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    value_age = []

    def value_pass(n_steps, t_first):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps + 1)]
        launches0 = ctx.launches
        copies0 = b.resample_stats()[0]
        value_age.append(age[0])
        w0 = time.time()
        ev[0].record(stream)
        for t in range(n_steps):
            step(t_first + t)
            ev[t + 1].record(stream)
        barrier()
        sampler.window(w0, time.time())
        per = [ev[t].elapsed_time(ev[t + 1]) for t in range(n_steps)]
        ms = max_over_ranks(ev[0].elapsed_time(ev[n_steps]))
        return ms, per, ctx.launches - launches0, b.resample_stats()[0] - copies0

    def kernel_pass(n_steps, t_first):
        barrier()
        age0 = age[0]
        copies0 = b.resample_stats()[0]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.profile_begin()
        ev0.record(stream)
        for t in range(n_steps):
            step(t_first + t)
        ev1.record(stream)
        barrier()
        ctx.profile_end()
        ms = ev0.elapsed_time(ev1)
        copies = b.resample_stats()[0] - copies0
        table = {}
        for name, (kms, cnt) in ctx.kernel_times().items():
            per_launch = kms / max(cnt, 1)
            if name.startswith("k_gather"):
                alg = bytes_per_particle * n_local
            elif name.startswith("k_copy_inplace") and journal_updates:
                jp = 4 * ((J_nodes + 1 + 3) // 4 * 4)
                alg = 2 * (16 + jp * (age0 + 1 + (n_steps - 1) / 2.0) + 12) * copies / max(cnt, 1)
            elif name.startswith("k_copy_inplace"):
                alg = bytes_per_particle * copies / max(cnt, 1)
            elif name.startswith("k_propose_journal"):
                # per particle: its journal read once (16 B header + 48 B per earlier update, averaged over the
                # launches of this pass), 48 B appended, state r+w, weight r+w, prototype id
                jp = 4 * ((J_nodes + 1 + 3) // 4 * 4)
                alg = (16 + jp * (age0 + (n_steps - 1) / 2.0) + jp + 8 + 16 + 4) * n_local
            elif name.startswith("k_propose"):
                alg = bytes_propose * n_local
            else:
                alg = None
            table[name] = {"ms_per_launch": round(per_launch, 5), "launches": cnt,
                           "share_of_step": round(kms / ms, 4),
                           "algorithmic_GBps": None if alg is None else round(alg / (per_launch * 1e-3) / 1e9, 1)}
        return ms, table, copies

    try:
        ncu_ratio = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except (OSError, ValueError):
        ncu_ratio = {}

    def roofline_of(table, name_prefix, note):
        name = next((k for k in table if k.startswith(name_prefix)), None)
        if name is None:
            return None
        row = table[name]
        ach = row["algorithmic_GBps"] or 0.0
        alg_bytes = ach * 1e9 * row["ms_per_launch"] * 1e-3
        ratio = ncu_ratio.get(name_prefix, {}).get("traffic_over_algorithmic") if wl == "sysadmin" else None
        return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "algorithmic_bytes_per_launch": alg_bytes,
                "traffic": None if ratio is None else alg_bytes * ratio,
                "traffic_source": ncu_ratio.get(name_prefix, {}).get("source") if ratio is not None else None,
                "peak_source": peak_src, "kernel_ms": row["ms_per_launch"],
                "share_of_step": row["share_of_step"], "note": note}

    # ---- pass 1 + 2: the default path, in-place systematic resampling (survivors are not moved) ----
    ms, per_step, launches, copies = value_pass(steps, args.warmup + 2)
    value_age0 = value_age[0]
    kms_total, table, kcopies = kernel_pass(min(steps, 10), 500)
    value = n_local * world * steps / (ms * 1e-3)
    dominant = max(table, key=lambda k: table[k]["ms_per_launch"] * table[k]["launches"])
    copied_frac = copies / float(n_local * steps)
    if dominant.startswith("k_copy_inplace"):
        roof = roofline_of(table, "k_copy_inplace",
                           "bytes = %d B x copied particles (%.1f%% of the particles per update; survivors "
                           "stay in place)" % (bytes_per_particle, 100 * kcopies / float(n_local * min(steps, 10))))
    elif journal_updates:
        roof = roofline_of(table, "k_propose_journal",
                           "bytes = each particle's journal read once (16 B + 48 B per earlier update; updates %d-%d "
                           "in this pass) + 48 B appended + state / weight" % (age[0] - min(steps, 10), age[0] - 1))
    else:
        roof = roofline_of(table, "k_propose",
                           "k_propose touches %d algorithmic bytes per particle in %d scattered rows; it is "
                           "bound by 32-byte-sector random access, not by streaming bandwidth"
                           % (bytes_propose, FS + FO))
    # whole update against the roofline: algorithmic bytes of propose + copies per device-second
    step_alg = bytes_propose * n_local + bytes_per_particle * copies / float(steps)
    if journal_updates:  # journal read + append per particle, copies move journals: averaged over the value pass
        jp = 4 * ((J_nodes + 1 + 3) // 4 * 4)
        mean_len = value_age0 + (steps - 1) / 2.0
        step_alg = (16 + jp * mean_len + jp + 28) * n_local + 2 * (16 + jp * mean_len) * copies / float(steps)
    step_frac = step_alg / (ms / steps * 1e-3) / 1e9 / peak

    # ---- the full-copy path (every particle gathered into the second buffer): the algorithm SURVEY.md
    #      §8d's B_upd describes and the north star's ">= 60 % of HBM peak" refers to ----
    full = None
    if headline and world == 1 and not args.no_full_copy and not sharded:
        ctx.set_option("inplace_resample", 0)
        for t in range(3):
            step(t)
        fms, ftable, _ = kernel_pass(min(steps, 10), 100)
        ctx.set_option("inplace_resample", 1)
        full = {"value": n_local * min(steps, 10) / (fms * 1e-3), "unit": UNIT,
                "ms_per_step": fms / min(steps, 10),
                "roofline": roofline_of(ftable, "k_gather", "bytes = %d B x every particle" % bytes_per_particle)}

    # ---- pass 3: end to end ----
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    lik = None
    for t in range(steps):
        lik = step(args.warmup + steps + t, likelihood=True)
        d2h = 8
        u_owner = float(shared.random_sample())
        if sharded:
            owner, i = b.sample_global(rng, u_owner)
        else:
            owner, i = 0, b.sample(rng)
        if i is not None:
            part = b.download(i, 1)
            d2h += part["counts"].nbytes + 4 + 4 + 8 + 4
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_local * world * steps / e2e_s
    d2h = int(max_over_ranks(float(d2h)))
    timeouts = b.timeouts() if sharded else 0
    phases = dict(getattr(b, "phase_ms", {}))
    exchange = getattr(b, "exchange", None)
    b.free()
    sim.close()

    res = {
        "value": value, "ms_per_step": ms / steps,
        "ms_per_step_spread_rank0": {"min": float(np.min(per_step)), "median": float(np.median(per_step)),
                                     "max": float(np.max(per_step))},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / steps * 1e3, "last_global_likelihood": lik},
        "gpu_launches": int(launches), "roofline": roof, "kernels": table,
        "step_roofline": {"algorithmic_bytes_per_step_per_gpu": step_alg, "achieved": step_alg / (ms / steps * 1e-3) / 1e9,
                          "peak": peak, "unit": "GB/s", "frac": step_frac,
                          "note": "propose + in-place copies, algorithmic bytes per GPU / device time per update"},
        "resampling_copies_per_update_frac": copied_frac, "full_copy": full, "p2p_timeouts": timeouts,
        "updates_per_particle_during_value_pass": [value_age0, value_age0 + steps - 1],
        "config": {"workload": WORKLOADS[wl]["text"], "structures": int(len(psid)),
                   "particles_per_gpu": n_local, "particles_total": n_local * world,
                   "count_cells_per_particle": C, "rng": "philox4x32-10",
                   "resampling": "systematic, in place (survivors keep their slot; duplicates fill dead slots); "
                                 "statistical parity (PHILOX); bit-exact replay parity is the full-copy REPLAY path",
                   "algorithmic_bytes_per_particle_copy": bytes_per_particle,
                   "algorithmic_bytes_per_particle_propose": bytes_propose,
                   "parallelism": "particles sharded, %d rank(s)%s" % (
                       world, ", exchange=%s (no NCCL on the update path)" % exchange if sharded else ""),
                   "last_step_host_ms_rank0": phases,
                   "l2": "inputs (%.1f GB of counts per GPU) exceed the 126 MB L2; no flush needed"
                         % (n_local * C * 4 / 1e9),
                   "passes": "value: K async updates, one CUDA event per step; kernels: separate pass with an event "
                             "pair per launch; e2e: separate pass, host args in / host results out",
                   "e2e_note": "per step: (a,o) in as call arguments, GLOBAL likelihood (8 B) out on every rank, "
                               "then Belief::sample() + download of that particle on the rank that owns it"},
    }
    return res


def main_ours(args):
    import torch
    import torch.distributed as dist
    import fba_pomdp_b200 as fba

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = fba.Context(local_rank)
    sampler = ClockSampler(local_rank)

    head = belief_leg(args.workload, args, fba, torch, dist, ctx, world, rank, sampler, args.steps, True)
    extra = None
    if args.workload == "sysadmin" and not args.no_config4 and not args.particles:
        # BASELINE.json configs[3] in the same driver-run record: collision avoidance, 10^6 particles per GPU
        extra = belief_leg("ca", args, fba, torch, dist, ctx, world, rank, sampler, args.steps, False)
    journal = None
    if args.workload == "sysadmin" and not args.no_journal and not args.particles and not args.journal_updates:
        # the same workload with base+journal storage (young beliefs: particle = prototype + its increments)
        journal = belief_leg("sysadmin", args, fba, torch, dist, ctx, world, rank, sampler, args.steps, False,
                             journal_updates=96)
    clocks = sampler.finish()

    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 counts / f64 weights", "data": "synthetic",
        "config": head["config"], "clocks": clocks[0] if clocks else None,
        "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
        "ms_per_step_spread_rank0": head["ms_per_step_spread_rank0"], "step_roofline": head["step_roofline"],
        "kernels": head["kernels"], "resampling_copies_per_update_frac": head["resampling_copies_per_update_frac"],
        "full_copy": head["full_copy"], "p2p_timeouts": head["p2p_timeouts"],
    }
    if head["e2e"]["value"] > head["value"]:
        line["e2e"]["note"] = "e2e exceeds the device-timed value: the two passes ran different script steps"
    if extra is not None:
        line["config4_ca"] = {k: extra[k] for k in ("value", "ms_per_step", "ms_per_step_spread_rank0", "e2e",
                                                    "roofline", "step_roofline", "resampling_copies_per_update_frac",
                                                    "p2p_timeouts")}
        line["config4_ca"].update(unit=UNIT, n_gpus=world, workload=extra["config"]["workload"],
                                  particles_per_gpu=extra["config"]["particles_per_gpu"],
                                  structures=extra["config"]["structures"],
                                  clocks=clocks[1] if len(clocks) > 1 else None)
    if journal is not None:
        line["journal"] = {k: journal[k] for k in ("value", "ms_per_step", "ms_per_step_spread_rank0", "e2e", "roofline",
                                                   "step_roofline", "kernels", "resampling_copies_per_update_frac",
                                                   "updates_per_particle_during_value_pass", "p2p_timeouts")}
        line["journal"].update(
            unit=UNIT, n_gpus=world,
            storage="base + journal (fba_model_desc.delta_capacity = 11 x 96): a particle is its prior prototype + the "
                    "cells it incremented; valid for the first 96 updates of a belief, bit-identical results "
                    "(tests/test_cuda_journal.py). The headline `value` is the dense steady state, valid at any age.",
            clocks=clocks[2 if extra is not None else 1] if len(clocks) > (2 if extra is not None else 1) else None)
    if not args.no_rollouts:
        line["rollouts"] = rollouts_leg(ctx, fba, args, torch, world, rank, dist)
    if world == 1 and not args.no_rollouts:
        line["many_runs"] = many_runs_leg(ctx, fba, args)
        line["structure_beliefs"] = structure_beliefs_leg(ctx, fba, args)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = args.ref_particles
        v, s_per_step, kind = run_reference_cpu(n, 6, 1, 1, args.workload)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "the reference's BAImportanceSampling::updateEstimation, 1 thread (it has no "
                      "threads), %d particles (its resample is O(N^2)), 6 steps after 1 warm-up; "
                      "host has %d cores" % (n, os.cpu_count() or 1)}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sysadmin", choices=sorted(WORKLOADS),
                    help="sysadmin = BASELINE configs[4] (default, the headline); ca = configs[3]")
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (0: the workload's own)")
    ap.add_argument("--ref-particles", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rollouts", action="store_true", help="skip the POMCP rollouts leg (config 3)")
    ap.add_argument("--rollout-particles", type=int, default=1_000_000)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "allgather"],
                    help="how sharded beliefs ship surplus particles between GPUs")
    ap.add_argument("--force-sharded", action="store_true", help="use the sharded code path on 1 GPU")
    ap.add_argument("--no-full-copy", action="store_true", help="skip the full-copy resampling leg")
    ap.add_argument("--no-config4", action="store_true", help="skip the collision-avoidance (configs[3]) leg")
    ap.add_argument("--no-journal", action="store_true", help="skip the base+journal storage leg")
    ap.add_argument("--journal-updates", type=int, default=0,
                    help="> 0: base+journal storage holding this many updates per particle (0: dense blocks)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
