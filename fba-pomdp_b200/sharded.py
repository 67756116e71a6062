"""Particles sharded across the GPUs of one box: one process per GPU, `torch.distributed` for the
plumbing (NCCL on GPUs, gloo in the CPU tests of the host logic).

Per belief update the ranks exchange:
  * an all-gather of one double per rank — the shard's un-normalised weight total — so every rank
    forms the same global total and the same offspring quotas;
  * particles only where resampling leaves a rank over / under its capacity: the surplus offspring
    of over-quota ranks are shipped (all-to-all-v of particle records) into the empty slots of
    under-quota ranks. With balanced weight shares that is O(sqrt(N/G)) particles per update.
No collective touches the count blocks otherwise: the data path stays local to each GPU's HBM.
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

    def __init__(self, n_local, group=None):
        super().__init__(n_local)
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.moved_last = 0
        self.phase_ms = {}
        self._bufs = None

    def rank_rng(self, seed):
        """A PHILOX source whose key differs per rank, so shards draw independent streams."""
        mixed = (int(seed) + (self.rank + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        return capi.Rng.philox(mixed)

    def _setup(self):
        import torch
        L, h = self.L, self.h
        self._stream = torch.cuda.ExternalStream(self.ctx.stream)
        scal = L.fba_belief_scalars_ptr(h)
        self._local = torch.as_tensor(_RawCuda(scal, 8, "<f8", 8), device="cuda")  # view of scal[0]
        self._totals = torch.empty(self.world, dtype=torch.float64, device="cuda")
        self._totals_host = torch.empty(self.world, dtype=torch.float64).pin_memory()
        self._plan = np.zeros((self.world, self.world), np.int64)
        self._event = torch.cuda.Event()
        # export staging for the surplus of an over-quota shard: 1/64 of the shard by default
        _check(self.ctx.h, L.fba_belief_reserve_export(h, max(1024, self._n // 64)))
        if self.world > 1:
            # establish every pairwise NCCL connection now (the exchange plan picks different pairs
            # from step to step; first use of a pair costs milliseconds)
            rb = L.fba_belief_record_bytes(h)
            warm_plan = np.ones((self.world, self.world), np.int64) - np.eye(self.world, dtype=np.int64)
            src = torch.zeros((self.world - 1) * rb, dtype=torch.uint8, device="cuda")
            exchange_records(self.dist, self.group, warm_plan, self.rank, src, rb)
            self.dist.all_gather_into_tensor(self._totals, torch.zeros(1, dtype=torch.float64, device="cuda"),
                                             group=self.group)
            torch.cuda.synchronize()
        self._bufs = True

    def free(self, _simulator=None):
        """Drops the torch views / buffers tied to the library's stream before the belief (and later
        the context that owns the stream) goes away."""
        import torch
        if self._bufs:
            torch.cuda.synchronize()
            self._local = self._totals = self._totals_host = self._stream = self._event = None
            self._bufs = None
        super().free()

    def updateEstimation(self, a, o, rng, step_uniform=0.5):
        """One global importance-sampling update + resample. `step_uniform` in [0,1) must be the
        same on every rank (the shared systematic offset of the quota allocation).

        Everything is enqueued on the library's stream without waiting for the GPU; the host only
        waits for the G shard totals (a D2H copy ordered BEFORE the resampling kernels), which it
        needs for the all-to-all's split sizes, while the GPU is already resampling."""
        import time
        import torch
        L, h, ctx = self.L, self.h, self.ctx
        if self._bufs is None:
            self._setup()
        u = float(step_uniform)
        t0 = time.perf_counter()
        with torch.cuda.stream(self._stream):
            # phase 1: step + weights + shard total (left on the device)
            _check(ctx.h, L.fba_belief_propose(h, a, o, C.byref(rng), None))
            if self.world > 1:
                self.dist.all_gather_into_tensor(self._totals, self._local, group=self.group)
            else:
                self._totals.copy_(self._local)
            self._totals_host.copy_(self._totals, non_blocking=True)
            self._event.record(self._stream)
            # phases 2-3: quota on device, normalise, resample in place, surplus -> export buffer
            _check(ctx.h, L.fba_belief_shard_resample_async(h, self._totals.data_ptr(), self.world, self.rank,
                                                            u, C.byref(rng)))
            self._event.synchronize()  # totals are on the host; the GPU keeps resampling
            t1 = time.perf_counter()
            plan, tot = self._plan, C.c_double(0)
            _check(ctx.h, L.fba_belief_shard_plan(h, self._totals_host.data_ptr(), self.world, self.rank, u,
                                                  plan.ctypes.data_as(C.c_void_p), C.byref(tot)))
            self.moved_last = int(plan.sum())
            if self.moved_last:
                rb = L.fba_belief_record_bytes(h)
                n_out, n_in = int(plan[self.rank].sum()), int(plan[:, self.rank].sum())
                src = (torch.as_tensor(_RawCuda(L.fba_belief_export_ptr(h), n_out * rb), device="cuda")
                       if n_out else torch.empty(0, dtype=torch.uint8, device="cuda"))
                dst = (torch.as_tensor(_RawCuda(L.fba_belief_import_ptr(h, n_in), n_in * rb), device="cuda")
                       if n_in else torch.empty(0, dtype=torch.uint8, device="cuda"))
                exchange_records(self.dist, self.group, plan, self.rank, src, rb, dst)
                # phase 4: imported records fill the slots the local resample left dead
                _check(ctx.h, L.fba_belief_import(h, n_in))
        self.phase_ms = {"enqueue propose..resample + wait for totals": (t1 - t0) * 1e3,
                         "plan + exchange (enqueue)": (time.perf_counter() - t1) * 1e3}
        return tot.value
