#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generates tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref/libfba_ref.so, built by `make -C oracle ref` from /root/reference) under seed "42".

Each fixture holds, for one configuration of BASELINE.json:
  * the model description in this repo's vocabulary (sizes, features, domain functor, start sampler),
  * the structure table + initial particles the reference's prior produced,
  * an (action, observation) script from the reference's true environment under a random policy,
  * per operation: the exact mt19937 words the reference consumed (the replay stream) and the
    reference's results (domain states, weights, per-particle count sums, final full counts).

Run from the repo root:  python oracle/gen_golden.py [name ...]
The fixtures are committed; /root/reference is not needed to *use* them.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pyref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# name -> reference flags (-D, --size, --width, --height, factored?, --structure-prior) and sizes
CONFIGS = {
    # BASELINE.json configs[0]: episodic-tiger tabular BA-POMDP, IS belief, 1024 particles
    "tiger": dict(domain="episodic-tiger", factored=False, N=1024, steps=20, rs_N=256),
    # configs[1]: factored tiger, 8 irrelevant features (reduced N; reinvigoration fixture below)
    "ftiger": dict(domain="episodic-factored-tiger", size=8, factored=True, N=256, steps=20, rs_N=128),
    "ftiger_mu": dict(domain="episodic-factored-tiger", size=8, factored=True,
                      structure_prior="match-uniform", N=256, steps=12, reinv_N=96, reinv_K=24),
    # configs[2]: gridworld (size 3, the size the reference's own tests use, test/test.cpp:106-107)
    "gridworld3": dict(domain="gridworld", size=3, factored=False, N=48, steps=20, rs_N=0),
    # configs[3]: collision avoidance 5x5, 1 obstacle, heterogeneous structures
    "ca": dict(domain="centered-collision-avoidance", size=1, width=5, height=5, factored=True,
               structure_prior="match-uniform", N=256, steps=20, reinv_N=64, reinv_K=16),
    # configs[4]: linear sysadmin, 10 computers
    "sysadmin": dict(domain="linear-sysadmin", size=10, factored=True, N=96, steps=20, rs_N=48),
    # extra pins beyond the five configs:
    # sysadmin's `mutate` (chained subscripts drawn right to left) through reinvigoration
    "sysadmin3": dict(domain="linear-sysadmin", size=3, factored=True, N=64, steps=12, reinv_N=48, reinv_K=12),
    # gridworld size 5 (S = O = 150, 180 000 count cells = 720 KB per particle): the case that needs
    # base+delta storage on the GPU (SURVEY.md §7 hard part 4); the reference runs it with COW rows
    "gridworld5": dict(domain="gridworld", size=5, factored=False, N=24, steps=20, rs_N=16),
    # factored gridworld: three OBSERVATION features (likelihood = double product of float factors)
    "gridworld3_fba": dict(domain="gridworld", size=3, factored=True, N=64, steps=20),
}

HORIZON = 20
DISCOUNT = 0.95


def model_desc(r, cfg):
    """Model description dict from the live reference objects."""
    L = r.L
    import ctypes as C
    oi = np.zeros(40, np.int32)
    od = np.zeros(9, np.float64)
    sv = np.zeros(max(r.S, 1), np.float32)
    stab = np.zeros(64, np.int32)
    L.ref_domain_desc.argtypes = [C.c_void_p] * 5
    L.ref_domain_desc(r.h, pyref._p(oi), pyref._p(od), pyref._p(sv), pyref._p(stab))
    assert oi[0] >= 0, "domain not described"
    d = dict(S=r.S, A=r.A, O=r.O, feat_s=r.feat_s.copy(), feat_o=r.feat_o.copy(),
             tabular=int(not cfg["factored"]), domain=int(oi[0]), action_draw=int(oi[1]),
             start_kind=int(oi[2]), start_ip=oi[3:7].copy(), dom_ip=oi[8:40].copy(),
             dom_dp=od[:8].copy(), start_total=float(od[8]),
             start_values=sv if oi[2] == 4 else np.zeros(0, np.float32),
             start_table=stab if oi[2] == 3 else np.zeros(0, np.int32))
    return d


class StructTable:
    def __init__(self):
        self.keys = {}
        self.t, self.o = [], []

    def add(self, tp, op):
        k = (tp.tobytes(), op.tobytes())
        if k not in self.keys:
            self.keys[k] = len(self.t)
            self.t.append(tp.reshape(-1).copy())
            self.o.append(op.reshape(-1).copy())
        return self.keys[k]


def dump_filter(r, filt, table, stride=None):
    """(struct_id[N], states[N], counts[N,stride]) of a reference filter."""
    n = r.size(filt)
    parts = [r.particle(filt, i) for i in range(n)]
    sid = np.array([table.add(tp, op) for tp, op, _ in parts], np.int32)
    if stride is None:
        stride = max(len(c) for _, _, c in parts)
    counts = np.zeros((n, stride), np.float32)
    for i, (_, _, c) in enumerate(parts):
        counts[i, :len(c)] = c
    return sid, r.states(filt), counts


def count_sums(counts):
    return counts.astype(np.float64).sum(axis=1)


def gen(name, cfg):
    kw = dict(size=cfg.get("size", 0), width=cfg.get("width", 0), height=cfg.get("height", 0),
              factored=cfg["factored"], structure_prior=cfg.get("structure_prior", ""),
              discount=DISCOUNT, horizon=HORIZON, seed="42")
    r = pyref.Ref(cfg["domain"], **kw)
    out = {}
    desc = model_desc(r, cfg)
    for k, v in desc.items():
        out["model/" + k] = np.asarray(v)
    out["meta/discount"] = np.float64(DISCOUNT)
    out["meta/horizon"] = np.int32(HORIZON)

    # (a,o) script from the true environment, random policy (SURVEY.md §8d "synthetic episodes")
    acts, obs, flags = r.env_script(cfg["steps"], HORIZON)
    out["script/a"], out["script/o"], out["script/flags"] = acts, obs, flags

    # reward / terminal samples for the domain functor
    rs = np.random.RandomState(7)
    trip = np.stack([rs.randint(0, r.S, 4000), rs.randint(0, r.A, 4000), rs.randint(0, r.S, 4000)], 1)
    if r.S * r.A * r.S <= 4000:
        trip = np.array([(s, a, s2) for s in range(r.S) for a in range(r.A) for s2 in range(r.S)])
    rew = np.array([r.reward(int(s), int(a), int(s2)) for s, a, s2 in trip])
    out["functor/triples"] = trip.astype(np.int32)
    out["functor/reward"] = rew[:, 0].astype(np.float64)
    out["functor/terminal"] = rew[:, 1].astype(np.uint8)

    table = StructTable()

    # ---------------- importance sampling ----------------
    N = cfg["N"]
    r.reseed("42")
    r.mark()
    r.belief_init(pyref.F_IS, N)
    out["is/init_words"] = r.words_since_mark()
    sid, st, counts = dump_filter(r, pyref.F_IS, table)
    # heterogeneous structures: stride must cover every structure seen later too; the IS path never
    # creates structures, so the max over the initial particles is enough.
    stride = counts.shape[1]
    out["is/init_struct_id"], out["is/init_state"], out["is/init_counts"] = sid, st, counts

    n_updates = 0
    for t in range(cfg["steps"]):
        a, o, fl = int(acts[t]), int(obs[t]), int(flags[t])
        if fl & 2 and t > 0:
            # new episode: BAPOMDPExperiment.cpp:59 resetDomainStateDistribution
            r.mark()
            r.reset_domain_states(pyref.F_IS)
            out["is/%d/reset_words" % t] = r.words_since_mark()
            out["is/%d/reset_state" % t] = r.states(pyref.F_IS)
            _, _, c = dump_filter(r, pyref.F_IS, table, stride)
            out["is/%d/reset_count_sums" % t] = count_sums(c)
        if fl & 1:
            continue  # terminal step: Episode.cpp:47-50 skips the belief update
        r.mark()
        total = r.is_update(a, o)
        out["is/%d/update_words" % t] = r.words_since_mark()
        w, tw = r.is_weights()
        out["is/%d/likelihood" % t] = np.float64(total)
        out["is/%d/w" % t] = w
        out["is/%d/total_weight" % t] = np.float64(tw)
        out["is/%d/state" % t] = r.states(pyref.F_IS)
        _, _, c = dump_filter(r, pyref.F_IS, table, stride)
        out["is/%d/count_sums" % t] = count_sums(c)
        if n_updates == 0:
            out["is/%d/counts" % t] = c
        r.mark()
        r.is_resample()
        out["is/%d/resample_words" % t] = r.words_since_mark()
        out["is/%d/rs_state" % t] = r.states(pyref.F_IS)
        sid2, _, c = dump_filter(r, pyref.F_IS, table, stride)
        out["is/%d/rs_struct_id" % t] = sid2
        out["is/%d/rs_count_sums" % t] = count_sums(c)
        w, tw = r.is_weights()
        out["is/%d/rs_total_weight" % t] = np.float64(tw)
        n_updates += 1
        last = t
    sid, st, counts = dump_filter(r, pyref.F_IS, table, stride)
    out["is/final_struct_id"], out["is/final_state"], out["is/final_counts"] = sid, st, counts
    out["is/last_step"] = np.int32(last)

    # ---------------- rollouts on the final IS belief ----------------
    n_roll = 64
    rs = np.random.RandomState(11)
    pid = rs.randint(0, N, n_roll).astype(np.int32)
    start = rs.randint(0, r.S, n_roll).astype(np.int32)
    depth = rs.randint(1, HORIZON + 1, n_roll).astype(np.int32)
    rets, words, offs = [], [], [0]
    for i in range(n_roll):
        r.mark()
        rets.append(r.rollout(pyref.F_IS, int(pid[i]), int(start[i]), int(depth[i])))
        wds = r.words_since_mark()
        words.append(wds)
        offs.append(offs[-1] + len(wds))
    out["roll/particle"], out["roll/start"], out["roll/depth"] = pid, start, depth
    out["roll/ret"] = np.array(rets, np.float64)
    out["roll/words"] = np.concatenate(words) if words else np.zeros(0, np.uint32)
    out["roll/offsets"] = np.array(offs, np.int64)

    # ---------------- rejection sampling ----------------
    if cfg.get("rs_N"):
        M = cfg["rs_N"]
        r.reseed("43")
        r.mark()
        r.belief_init(pyref.F_RS, M)
        out["rs/init_words"] = r.words_since_mark()
        sid, st, counts = dump_filter(r, pyref.F_RS, table, stride)
        out["rs/init_struct_id"], out["rs/init_state"], out["rs/init_counts"] = sid, st, counts
        done = 0
        for t in range(cfg["steps"]):
            a, o, fl = int(acts[t]), int(obs[t]), int(flags[t])
            if fl & 2 and t > 0:
                r.mark()
                r.reset_domain_states(pyref.F_RS)
                out["rs/%d/reset_words" % t] = r.words_since_mark()
                out["rs/%d/reset_state" % t] = r.states(pyref.F_RS)
            if fl & 1:
                continue
            r.mark()
            r.update_estimation(pyref.F_RS, a, o)
            out["rs/%d/words" % t] = r.words_since_mark()
            sid, st, c = dump_filter(r, pyref.F_RS, table, stride)
            out["rs/%d/state" % t] = st
            out["rs/%d/struct_id" % t] = sid
            out["rs/%d/count_sums" % t] = count_sums(c)
            done += 1
            if done >= 6:
                break
        out["rs/last_step"] = np.int32(t)
        sid, st, counts = dump_filter(r, pyref.F_RS, table, stride)
        out["rs/final_struct_id"], out["rs/final_state"], out["rs/final_counts"] = sid, st, counts

    # ---------------- reinvigoration (two flat filters) ----------------
    if cfg.get("reinv_N"):
        M, K = cfg["reinv_N"], cfg["reinv_K"]
        r.reseed("44")
        r.mark()
        r.belief_init(pyref.F_REINV, M, K)
        out["reinv/init_words"] = r.words_since_mark()
        out["reinv/K"] = np.int32(K)
        # the fully connected filter bounds the stride
        fc_parts = [r.particle(pyref.F_REINV_FC, i) for i in range(M)]
        rstride = max(len(c) for _, _, c in fc_parts)
        out["reinv/stride"] = np.int64(rstride)
        for tag, filt in (("b", pyref.F_REINV), ("fc", pyref.F_REINV_FC)):
            sid, st, counts = dump_filter(r, filt, table, rstride)
            out["reinv/init_%s_struct_id" % tag] = sid
            out["reinv/init_%s_state" % tag] = st
            out["reinv/init_%s_counts" % tag] = counts
        done = 0
        for t in range(cfg["steps"]):
            a, o, fl = int(acts[t]), int(obs[t]), int(flags[t])
            if fl & 2 and t > 0:
                r.mark()
                r.reset_domain_states(pyref.F_REINV)
                out["reinv/%d/reset_words" % t] = r.words_since_mark()
                out["reinv/%d/reset_b_state" % t] = r.states(pyref.F_REINV)
                out["reinv/%d/reset_fc_state" % t] = r.states(pyref.F_REINV_FC)
            if fl & 1:
                continue
            # split: reinvigorateParticles, then the two rejectSample calls (the same draws as
            # updateEstimation, ReinvigoratingRejectionSampling.cpp:89-106)
            r.mark()
            r.reinvigorate_only()
            out["reinv/%d/breed_words" % t] = r.words_since_mark()
            sid, st, c = dump_filter(r, pyref.F_REINV, table, rstride)
            out["reinv/%d/breed_b_struct_id" % t] = sid
            out["reinv/%d/breed_b_state" % t] = st
            out["reinv/%d/breed_b_counts" % t] = c
            # second half through the real updateEstimation would breed again, so replay the two
            # rejectSample calls directly on the private filters
            import ctypes as C
            r.L.ref_reinv_reject_only.argtypes = [C.c_void_p, C.c_int, C.c_int]
            r.mark()
            r.L.ref_reinv_reject_only(r.h, a, o)
            out["reinv/%d/reject_words" % t] = r.words_since_mark()
            for tag, filt in (("b", pyref.F_REINV), ("fc", pyref.F_REINV_FC)):
                sid, st, c = dump_filter(r, filt, table, rstride)
                out["reinv/%d/%s_struct_id" % (t, tag)] = sid
                out["reinv/%d/%s_state" % (t, tag)] = st
                out["reinv/%d/%s_count_sums" % (t, tag)] = count_sums(c)
            done += 1
            if done >= 4:
                break
        out["reinv/last_step"] = np.int32(t)
        for tag, filt in (("b", pyref.F_REINV), ("fc", pyref.F_REINV_FC)):
            sid, st, counts = dump_filter(r, filt, table, rstride)
            out["reinv/final_%s_struct_id" % tag] = sid
            out["reinv/final_%s_state" % tag] = st
            out["reinv/final_%s_counts" % tag] = counts

    out["structs/t_par"] = np.array(table.t, np.uint32)
    out["structs/o_par"] = np.array(table.o, np.uint32)
    r.close()

    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-12s S=%d A=%d O=%d structs=%d stride=%d -> %s (%.1f KB)" % (
        name, desc["S"], desc["A"], desc["O"], len(table.t), stride, path,
        os.path.getsize(path) / 1024))


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        gen(n, CONFIGS[n])
