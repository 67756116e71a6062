#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generates tests/golden/gibbs.npz by running the UNMODIFIED reference's
MHwithinGibbs belief (src/beliefs/bayes-adaptive/factored/MHwithinGibbs.cpp) under seed "42" on
episodic-factored-tiger (3 irrelevant features, match-uniform structure prior), once per way of sampling the
state history (MSG = backward messages + forward sampling, RS = rejection sampling):

  * the belief after a few episodes of updateEstimation = reinvigorate's input, and the recorded history,
  * the prior model of EVERY structure the domain's mutate can reach, the domain state prior,
  * the exact mt19937 words the private MHwithinGibbs::reinvigorate consumed, and the belief it produced.

Run from the repo root:  python oracle/gen_gibbs.py       (needs oracle/_ref/libfba_ref.so)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as GG  # noqa: E402
import pyref as R  # noqa: E402

N, SIZE = 24, 3


def drive(r, acts, obs, flags, rs):
    r.composite_init(R.F_GIBBS, N, 1 if rs else 0, -1e300)    # the threshold is never reached by updateEstimation
    lens, ha, ho = [0], [], []
    # reinvigorate runs right after an update in the reference (MHwithinGibbs.cpp:319-322): the history's last
    # episode is never empty (msgSampleStateHistory reads episode.back())
    last = max(t for t in range(len(acts)) if not (flags[t] & 1))
    for t in range(last + 1):
        if flags[t] & 2 and t > 0:
            r.composite_reset(R.F_GIBBS)
            if lens[-1]:
                lens.append(0)
        if flags[t] & 1:
            continue
        r.composite_update(R.F_GIBBS, int(acts[t]), int(obs[t]))
        lens[-1] += 1
        ha.append(int(acts[t])), ho.append(int(obs[t]))
    return lens, ha, ho


def main():
    cfg = dict(domain="episodic-factored-tiger", size=SIZE, factored=True, structure_prior="match-uniform")
    kw = dict(size=SIZE, factored=True, structure_prior="match-uniform", discount=GG.DISCOUNT, horizon=GG.HORIZON,
              seed="42")
    out = {}
    r = R.Ref(cfg["domain"], **kw)
    for k, v in GG.model_desc(r, cfg).items():
        out["model/" + k] = np.asarray(v)
    out["meta/discount"], out["meta/horizon"] = np.float64(GG.DISCOUNT), np.int32(GG.HORIZON)
    out["model_state_prior"] = r.state_prior()
    acts, obs, flags = r.env_script(36, GG.HORIZON)
    out["script/a"], out["script/o"], out["script/flags"] = acts, obs, flags
    table = GG.StructTable()
    for tag, rs in (("msg", False), ("rs", True)):
        r.reseed("48")
        lens, ha, ho = drive(r, acts, obs, flags, rs)
        out["history/len"], out["history/a"], out["history/o"] = (np.array(x, np.int32) for x in (lens, ha, ho))
        sid, st, counts = GG.dump_filter(r, R.F_GIBBS, table)
        if not rs:
            # every structure factored tiger's mutate reaches: any parent set of O[listen = 2][0]
            FS, FO = len(r.feat_s), len(r.feat_o)
            base_t, base_o = table.t[0].copy(), table.o[0].copy()
            for mask in range(1 << FS):
                o2 = base_o.copy()
                o2[2 * FO + 0] = mask
                table.add(base_t, o2)
            n_structs = len(table.t)
        w, tot = r.weights(R.F_GIBBS)
        out[tag + "/old_struct_id"], out[tag + "/old_state"], out[tag + "/old_counts"] = sid, st, counts
        out[tag + "/old_w"], out[tag + "/old_total_weight"] = w, np.float64(tot)
        r.mark()
        r.gibbs_run()
        out[tag + "/words"] = r.words_since_mark()
        assert r.gibbs_log_likelihood() == 0.0
        sid2, st2, c2 = GG.dump_filter(r, R.F_GIBBS, table)
        assert len(table.t) == n_structs, "reinvigorate produced a structure outside the enumerated table"
        out[tag + "/new_struct_id"], out[tag + "/new_state"], out[tag + "/new_counts"] = sid2, st2, c2
        print("%s: %d particles, %d history steps in %d episodes, reinvigorate consumed %d words, %d distinct structures after"
              % (tag, N, len(ha), len(lens), len(out[tag + "/words"]), len(np.unique(sid2))))
    priors = [r.prior_model(t, o) for t, o in zip(table.t, table.o)]
    stride = max(len(p) for p in priors)
    pc = np.zeros((len(priors), stride), np.float32)
    for k, p in enumerate(priors):
        pc[k, :len(p)] = p
    out["priors/counts"] = pc
    out["structs/t_par"], out["structs/o_par"] = np.stack(table.t), np.stack(table.o)
    np.savez_compressed(os.path.join(GG.OUT, "gibbs.npz"), **out)
    r.close()


if __name__ == "__main__":
    main()
