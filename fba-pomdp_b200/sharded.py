"""Particles sharded across the GPUs of one box: one process per GPU, `torch.distributed` for the
plumbing (NCCL on GPUs, gloo in the CPU tests of the host logic).

Per belief update the ranks exchange:
  * one double per rank — the shard's un-normalised weight total — so every rank forms the same
    global total and the same offspring quotas;
  * particles only where resampling leaves a rank over / under its capacity: the surplus offspring
    of over-quota ranks go into the empty slots of under-quota ranks. With balanced weight shares
    that is O(sqrt(N/G)) particles per update.
exchange="p2p" (default): both travel through peer-mapped memory inside the update's own kernels
(fba_belief_sharded_update: step-stamped flags, surplus blocks stored straight into the destination
GPU's dead slots over NVLink) — torch.distributed is used ONCE, at setup, to all-gather the CUDA IPC
handles; the update path has no collective and no host synchronisation.
exchange="allgather": NCCL all-gather of the totals + fixed-size windows of an export buffer, for
setups without peer mapping.
"""
import ctypes as C

import numpy as np

from . import capi
from .beliefs import BAImportanceSampling, _check


def offspring_quotas(shard_totals, n_total, u):
    """Systematic allocation of n_total offspring to shards in proportion to their weight totals:
    quota_g = #{j : (j + u) / n_total in (C_{g-1}, C_g]} with C the cumulative weight shares.
    Deterministic given (shard_totals, n_total, u), so every rank computes the same answer.
    Unbiased: E[quota_g] = n_total * W_g / W for u ~ U[0,1)."""
    w = np.asarray(shard_totals, np.float64)
    if not np.all(np.isfinite(w)) or w.sum() <= 0:
        raise capi.FbaError(capi.ERR_INVALID, "offspring_quotas: total weight must be positive")
    c = np.cumsum(w) / w.sum()
    c[-1] = 1.0
    edges = np.floor(c * n_total - u + 1.0).astype(np.int64)  # number of j with (j+u)/n <= C_g
    edges = np.clip(edges, 0, n_total)
    edges[-1] = n_total
    q = np.diff(np.concatenate([[0], edges]))
    return q.astype(np.int64)


def exchange_plan(quotas, capacity):
    """send[g][h] = particles rank g ships to rank h so that every rank ends with `capacity`
    particles. Greedy in rank order; identical on every rank."""
    q = np.asarray(quotas, np.int64)
    G = len(q)
    if q.sum() != capacity * G:
        raise capi.FbaError(capi.ERR_INVALID, "exchange_plan: quotas do not sum to the total capacity")
    surplus = np.maximum(q - capacity, 0)
    deficit = np.maximum(capacity - q, 0)
    send = np.zeros((G, G), np.int64)
    h = 0
    for g in range(G):
        while surplus[g] > 0:
            while deficit[h] == 0:
                h += 1
            k = min(surplus[g], deficit[h])
            send[g, h] += k
            surplus[g] -= k
            deficit[h] -= k
    return send


def exchange_records(dist, group, plan, rank, src, record_bytes, dst=None):
    """Ships particle records according to `plan` (exchange_plan): `src` holds this rank's surplus
    records (uint8, plan[rank].sum() * record_bytes), returns the records received. Works on CUDA
    tensors over NCCL and on CPU tensors over gloo."""
    import torch
    n_in = int(plan[:, rank].sum())
    if dst is None:
        dst = torch.empty(n_in * record_bytes, dtype=torch.uint8, device=src.device)
    dist.all_to_all_single(dst, src, [int(x) * record_bytes for x in plan[:, rank]],
                           [int(x) * record_bytes for x in plan[rank]], group=group)
    return dst


def split_requests(n_total, world):
    """Root-parallel split of n_total planner requests (leaf rollouts / simulations): rank g serves
    n_total // world of them, the first n_total % world ranks one more. Identical on every rank."""
    base, rem = divmod(int(n_total), int(world))
    return np.array([base + (1 if g < rem else 0) for g in range(world)], np.int64)


def gather_ragged(dist, group, local, counts, device="cpu"):
    """All ranks' result vectors (rank g contributes counts[g] doubles) concatenated in rank order, on
    every rank: one all-gather of fixed-size (max count) slots, 8 bytes per request (SURVEY.md §8e:
    "results returned by all-gather, 32 KB")."""
    import torch
    counts = np.asarray(counts, np.int64)
    world, slot = len(counts), int(counts.max()) if len(counts) else 0
    mine = torch.zeros(max(slot, 1), dtype=torch.float64, device=device)
    loc = torch.as_tensor(np.asarray(local, np.float64))
    mine[:len(loc)] = loc.to(device)
    if world == 1:
        return mine[:counts[0]].cpu().numpy()
    out = torch.empty(world * max(slot, 1), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    out = out.cpu().numpy().reshape(world, max(slot, 1))
    return np.concatenate([out[g, :counts[g]] for g in range(world)])


class _RawCuda:
    """Zero-copy view of a device pointer for torch.as_tensor."""

    def __init__(self, ptr, nbytes, typestr="|u1", itemsize=1):
        self.__cuda_array_interface__ = {"shape": (int(nbytes) // itemsize,), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3, "strides": None}


class ShardedBAImportanceSampling(BAImportanceSampling):
    """BAImportanceSampling over G shards of n_local particles (PHILOX mode). Rank-local work goes
    to the C ABI phases fba_belief_{propose,normalize,resample_shard,import}; NCCL collectives are
    enqueued on the library's own stream (made torch's current stream), so one update costs a
    single host synchronisation — the read-back of the G shard totals that the exchange plan
    (host-known split sizes for the all-to-all) needs."""

    def __init__(self, n_local, group=None, exchange="p2p"):
        """exchange: "p2p" — surplus records are stored by the resampling kernel directly into the
        destination GPU's import buffer over NVLink (CUDA IPC), plan on device, no host sync at all;
        "allgather" — fixed-size windows of the export buffers through one NCCL all-gather (for
        setups where peer mapping is unavailable)."""
        super().__init__(n_local)
        if exchange not in ("p2p", "allgather"):
            raise capi.FbaError(capi.ERR_INVALID, "exchange must be 'p2p' or 'allgather'")
        self.exchange = exchange
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.moved_last = 0
        self.phase_ms = {}
        self._bufs = None
        self.trace = None  # set to [] to record per-phase CUDA events (trace_summary)

    def _trace_events(self):
        import torch
        return [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    def trace_summary(self):
        """Mean device milliseconds per phase over the traced updates."""
        import torch
        torch.cuda.synchronize()
        names = ["propose+sums", "all_gather+D2H", "quota+normalise+resample", "exchange+import"]
        out = {}
        for k, nm in enumerate(names):
            out[nm] = float(np.mean([t[k].elapsed_time(t[k + 1]) for t in self.trace]))
        out["gap to next update"] = float(np.mean([a[4].elapsed_time(b[0])
                                                   for a, b in zip(self.trace[:-1], self.trace[1:])])) \
            if len(self.trace) > 1 else 0.0
        return out

    def rank_rng(self, seed):
        """A PHILOX source whose key differs per rank, so shards draw independent streams."""
        mixed = (int(seed) + (self.rank + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        return capi.Rng.philox(mixed)

    def _setup(self):
        import torch
        L, h = self.L, self.h
        if self.exchange == "p2p" and self._setup_p2p():
            self._bufs = True
            return
        self.exchange = "allgather"
        self._stream = torch.cuda.ExternalStream(self.ctx.stream)
        scal = L.fba_belief_scalars_ptr(h)
        self._local = torch.as_tensor(_RawCuda(scal, 8, "<f8", 8), device="cuda")  # view of scal[0]
        self._totals = torch.empty(self.world, dtype=torch.float64, device="cuda")
        self._totals_host = torch.empty(self.world, dtype=torch.float64).pin_memory()
        self._plan = np.zeros((self.world, self.world), np.int64)
        self._event = torch.cuda.Event()
        # export staging for the surplus of an over-quota shard: 1/64 of the shard by default (a
        # larger surplus raises FBA_ERR_CAPACITY on the over-quota rank); the exchange window is
        # capped so that the gathered buffer stays <= 1 GiB
        rb = L.fba_belief_record_bytes(h)
        cap = max(1024, self._n // 64)
        self._window_cap = 64
        while self._window_cap * 2 <= cap and self._window_cap * 2 * rb * self.world <= (1 << 30):
            self._window_cap *= 2
        _check(self.ctx.h, L.fba_belief_reserve_export(h, cap + self._window_cap))
        self._gather_buf = torch.empty(self.world * self._window_cap * rb, dtype=torch.uint8, device="cuda")
        if self.world > 1:
            w = 64
            while w <= self._window_cap:  # touch every window size once (NCCL algorithm selection)
                src = torch.as_tensor(_RawCuda(L.fba_belief_export_ptr(h), w * rb), device="cuda")
                self.dist.all_gather_into_tensor(self._gather_buf[:self.world * w * rb], src, group=self.group)
                w *= 2
            self.dist.all_gather_into_tensor(self._totals, torch.zeros(1, dtype=torch.float64, device="cuda"),
                                             group=self.group)
            torch.cuda.synchronize()
        self._bufs = True

    def _setup_p2p(self):
        """Publishes / maps the control blocks and particle arrays (CUDA IPC): the blobs travel through
        one all-gather, the only collective of this mode. If any rank cannot map its peers (no peer
        access / IPC in this container), every rank falls back to the all-gather exchange — the
        decision is made collectively. -> True when every rank is mapped."""
        import torch
        L, h = self.L, self.h
        ok = 1.0
        mine = np.zeros(capi.P2P_BLOB_BYTES, np.uint8)
        try:
            _check(self.ctx.h, L.fba_belief_p2p_export(h, mine.ctypes.data_as(C.c_void_p)))
        except capi.FbaError:
            ok = 0.0
        if self.world > 1:
            all_h = torch.empty(self.world * capi.P2P_BLOB_BYTES, dtype=torch.uint8, device="cuda")
            self.dist.all_gather_into_tensor(all_h, torch.from_numpy(mine).cuda(), group=self.group)
            blobs = all_h.cpu().numpy()
        else:
            blobs = mine
        blobs = np.ascontiguousarray(blobs)
        if ok:
            try:
                _check(self.ctx.h, L.fba_belief_p2p_open(h, blobs.ctypes.data_as(C.c_void_p), self.world, self.rank))
            except capi.FbaError:
                ok = 0.0
        if self.world > 1:
            flag = torch.tensor([ok], dtype=torch.float32, device="cuda")
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)  # also: every rank has mapped
            ok = float(flag.item())
        torch.cuda.synchronize()
        return bool(ok)

    def _exchange(self, plan):
        """Ships the surplus records: every rank contributes a fixed-size WINDOW of its export
        buffer to one all-gather (a ring collective whose connections exist since communicator
        setup — unlike send/recv pairs, nothing is connected lazily in the timed path), then imports
        the segments addressed to it. Window = the largest surplus, rounded up to a power of two;
        plans larger than the reserved staging run in several rounds."""
        import torch
        L, h, ctx = self.L, self.h, self.ctx
        rb = L.fba_belief_record_bytes(h)
        sent = plan.sum(1)
        window = 64
        while window < int(sent.max()):
            window *= 2
        window = min(window, self._window_cap)
        exp_ptr = L.fba_belief_export_ptr(h)
        rounds = -(-int(sent.max()) // window)
        for r in range(rounds):
            lo = r * window
            src = torch.as_tensor(_RawCuda(exp_ptr + lo * rb, window * rb), device="cuda")
            out = self._gather_buf[:self.world * window * rb]
            self.dist.all_gather_into_tensor(out, src, group=self.group)
            # sender g's records for rank h start at sum(plan[g, :h]) in its export buffer
            for g in range(self.world):
                n = int(plan[g, self.rank])
                if not n:
                    continue
                first = int(plan[g, :self.rank].sum())
                a, b_ = max(first, lo), min(first + n, lo + window)
                if a < b_:
                    addr = out.data_ptr() + (g * window + (a - lo)) * rb
                    _check(ctx.h, L.fba_belief_import_from(h, addr, b_ - a))

    def free(self, _simulator=None):
        """Drops the torch views / buffers tied to the library's stream before the belief (and later
        the context that owns the stream) goes away."""
        import torch
        if self._bufs:
            torch.cuda.synchronize()
            if self.world > 1 and self.exchange == "p2p":
                self.dist.barrier(group=self.group)  # peers may still be storing into this rank's arrays
            self._local = self._totals = self._totals_host = self._stream = self._event = self._gather_buf = None
            self._bufs = None
        super().free()

    def rollouts(self, n_total, depth, discount, rng, gather=True):
        """POMCP's leaf evaluation, root-parallel over the shards: the n_total
        random-policy rollouts (RBAPOUCT::rollout, RBAPOUCT.cpp:295-323) are split evenly over the
        ranks; each rank draws its share's root particles from ITS shard and runs them in one launch
        on its own GPU. After a global resample every shard holds the same number of equally weighted
        particles, so sampling inside the shard is sampling from the global belief. The only traffic
        is the all-gather of the returns (8 bytes each). -> all n_total returns (rank order) on every
        rank, or this rank's share if gather is False."""
        from .beliefs import rollouts as local_rollouts
        counts = split_requests(n_total, self.world)
        m = int(counts[self.rank])
        idx = np.zeros(m, np.int64)
        st = np.zeros(m, np.int32)
        if m:
            _check(self.ctx.h, self.L.fba_belief_sample_batch(self.h, C.byref(rng), m, idx.ctypes.data_as(C.c_void_p)))
            _check(self.ctx.h, self.L.fba_belief_gather_states(self.h, m, idx.ctypes.data_as(C.c_void_p),
                                                               st.ctypes.data_as(C.c_void_p)))
        ret = local_rollouts(self, idx, st, np.full(m, int(depth), np.int32), discount, rng) if m else np.zeros(0)
        if not gather:
            return ret
        import torch
        device = "cuda:%d" % torch.cuda.current_device() if (self.world > 1 and
                                                               self.dist.get_backend(self.group) == "nccl") else "cpu"
        return gather_ragged(self.dist, self.group, ret, counts, device)

    def updateEstimation(self, a, o, rng, step_uniform=0.5, likelihood=True):
        """One global importance-sampling update + resample (BAImportanceSampling::updateEstimation,
        BAImportanceSampling.cpp:74-88, over all shards). `step_uniform` in [0,1) must be the same on
        every rank (the shared systematic offset of the quota allocation). Returns the GLOBAL step
        likelihood total (sum of all shards' un-normalised weights; 8 bytes D2H, one stream
        synchronisation), or None with likelihood=False (nothing waits for the GPU).

        p2p: ONE C call enqueues the whole update; ranks meet inside the kernels through peer-mapped
        flags. allgather: the host waits for the G shard totals (a D2H copy ordered BEFORE the
        resampling kernels), which it needs for the exchange windows, while the GPU is already
        resampling."""
        import time
        import torch
        L, h, ctx = self.L, self.h, self.ctx
        if self._bufs is None:
            self._setup()
        u = float(step_uniform)
        t0 = time.perf_counter()
        if self.exchange == "p2p":
            tot = C.c_double(0)
            _check(ctx.h, L.fba_belief_sharded_update(h, a, o, C.byref(rng), u, C.byref(tot) if likelihood else None))
            self.phase_ms = {"enqueue whole update": (time.perf_counter() - t0) * 1e3}
            return tot.value if likelihood else None
        trace = self._trace_events() if self.trace is not None else None
        with torch.cuda.stream(self._stream):
            if trace:
                trace[0].record(self._stream)
            # phase 1: step + weights + shard total (left on the device)
            _check(ctx.h, L.fba_belief_propose(h, a, o, C.byref(rng), None))
            if trace:
                trace[1].record(self._stream)
            if self.world > 1:
                self.dist.all_gather_into_tensor(self._totals, self._local, group=self.group)
            else:
                self._totals.copy_(self._local)
            self._totals_host.copy_(self._totals, non_blocking=True)
            self._event.record(self._stream)
            if trace:
                trace[2].record(self._stream)
            # phases 2-3: quota on device, normalise, resample in place, surplus -> export buffer
            _check(ctx.h, L.fba_belief_shard_resample_async(h, self._totals.data_ptr(), self.world, self.rank,
                                                            u, C.byref(rng)))
            if trace:
                trace[3].record(self._stream)
            self._event.synchronize()  # totals are on the host; the GPU keeps resampling
            t1 = time.perf_counter()
            plan, tot = self._plan, C.c_double(0)
            _check(ctx.h, L.fba_belief_shard_plan(h, self._totals_host.data_ptr(), self.world, self.rank, u,
                                                  plan.ctypes.data_as(C.c_void_p), C.byref(tot)))
            self.moved_last = int(plan.sum())
            if self.moved_last:
                self._exchange(plan)
            if trace:
                trace[4].record(self._stream)
                self.trace.append(trace)
        self.phase_ms = {"enqueue propose..resample + wait for totals": (t1 - t0) * 1e3,
                         "plan + exchange (enqueue)": (time.perf_counter() - t1) * 1e3}
        return tot.value

    def sample_global(self, rng, shared_uniform):
        """Belief::sample() of the GLOBAL belief (WeightedFilter::sample, WeightedFilter.cpp:163-191):
        after a global resample every shard carries the same total weight, so the owning shard is
        floor(shared_uniform * world) — `shared_uniform` in [0,1) must be the same on every rank — and
        that rank draws inside its shard. -> (owner rank, local particle index or None on other ranks)."""
        owner = min(int(float(shared_uniform) * self.world), self.world - 1)
        return owner, (self.sample(rng) if owner == self.rank else None)

    def timeouts(self):
        """Cross-rank waits that were abandoned (p2p): 0 in a healthy run; anything else means a rank
        fell out of step and this belief is invalid."""
        return int(self.L.fba_belief_p2p_timeouts(self.h)) if self.exchange == "p2p" and self._bufs else 0
