#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generates tests/golden/mh.npz by running the UNMODIFIED reference's
MHNIPS2018 belief (src/beliefs/bayes-adaptive/factored/MHNIPS2018.cpp) under seed "42" on
episodic-factored-tiger (3 irrelevant features, match-uniform structure prior):

  * the belief after a few episodes of updateEstimation (particles, weights) = MH's input,
  * the (action, observation) history the belief recorded,
  * the prior model of EVERY structure the domain's mutate can reach (FBAPOMDPPrior::computePriorModel),
  * the exact mt19937 words the private MHNIPS2018::MH consumed, and the belief it produced.

Run from the repo root:  python oracle/gen_mh.py       (needs oracle/_ref/libfba_ref.so)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as GG  # noqa: E402
import pyref  # noqa: E402

F_MH = 4
N, SIZE = 48, 3


def drive(r, acts, obs, flags, threshold, stop_before_last=False):
    """initiate an MHNIPS2018 belief and feed it the script; returns (history lens, actions, observations,
    log likelihood after every update, index of the last update step)"""
    r.mh_init(N, threshold)
    upd = [t for t in range(len(acts)) if not (flags[t] & 1)]
    lens, ha, ho, ll = [0], [], [], []
    for t in range(len(acts)):
        if flags[t] & 2 and t > 0:    # episode start: resetDomainStateDistribution (new history episode)
            r.reset_domain_states(F_MH)
            if lens[-1]:
                lens.append(0)
        if flags[t] & 1:              # terminal step: no belief update (Episode.cpp:47-50)
            continue
        if stop_before_last and t == upd[-1]:
            return lens, ha, ho, ll, t
        r.update_estimation(F_MH, int(acts[t]), int(obs[t]))
        lens[-1] += 1
        ha.append(int(acts[t])), ho.append(int(obs[t]))
        ll.append(r.mh_log_likelihood())
    return lens, ha, ho, ll, upd[-1]


def main():
    cfg = dict(domain="episodic-factored-tiger", size=SIZE, factored=True, structure_prior="match-uniform")
    kw = dict(size=SIZE, factored=True, structure_prior="match-uniform", discount=GG.DISCOUNT, horizon=GG.HORIZON,
              seed="42")
    out = {}
    # the threshold is never reached by updateEstimation: the belief after the last update is MH's input
    r = pyref.Ref(cfg["domain"], **kw)
    for k, v in GG.model_desc(r, cfg).items():
        out["model/" + k] = np.asarray(v)
    out["meta/discount"], out["meta/horizon"] = np.float64(GG.DISCOUNT), np.int32(GG.HORIZON)
    acts, obs, flags = r.env_script(60, GG.HORIZON)
    out["script/a"], out["script/o"], out["script/flags"] = acts, obs, flags
    lens, ha, ho, ll, t_last = drive(r, acts, obs, flags, -1e300)
    out["history/len"], out["history/a"], out["history/o"] = (np.array(x, np.int32) for x in (lens, ha, ho))
    out["mh/log_likelihood_before"] = np.float64(ll[-1])
    table = GG.StructTable()
    sid, st, counts = GG.dump_filter(r, F_MH, table)
    # every structure factored tiger's mutate reaches: any parent set of O[listen = 2][0]
    FS, FO = len(r.feat_s), len(r.feat_o)
    base_t, base_o = table.t[0].copy(), table.o[0].copy()
    for mask in range(1 << FS):
        o2 = base_o.copy()
        o2[2 * FO + 0] = mask
        table.add(base_t, o2)
    out["old/struct_id"], out["old/state"], out["old/counts"] = sid, st, counts
    out["old/w"] = np.full(N, 1.0 / N)             # after resample (MHNIPS2018.cpp:171)
    priors = [r.prior_model(t, o) for t, o in zip(table.t, table.o)]
    stride = max(max(len(p) for p in priors), counts.shape[1])
    pc = np.zeros((len(priors), stride), np.float32)
    for k, p in enumerate(priors):
        pc[k, :len(p)] = p
    out["priors/counts"] = pc

    # the private MHNIPS2018::MH (MHNIPS2018.cpp:188-255) on that belief, its words tapped
    r.mark()
    r.mh_run()
    out["mh/words"] = r.words_since_mark()
    assert r.mh_log_likelihood() == 0.0
    table2 = GG.StructTable()
    table2.keys, table2.t, table2.o = dict(table.keys), list(table.t), list(table.o)
    sid2, st2, c2 = GG.dump_filter(r, F_MH, table2, stride)
    assert len(table2.t) == len(table.t), "MH produced a structure outside the enumerated table"
    out["new/struct_id"], out["new/state"], out["new/counts"] = sid2, st2, c2
    out["structs/t_par"], out["structs/o_par"] = np.stack(table.t), np.stack(table.o)
    np.savez_compressed(os.path.join(GG.OUT, "mh.npz"), **out)
    r.close()
    print("mh.npz: %d particles, %d history steps in %d episodes, %d structures, MH consumed %d words"
          % (N, len(ha), len(lens), len(table.t), len(out["mh/words"])))


if __name__ == "__main__":
    main()
