#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generates tests/golden/nested.npz by running the UNMODIFIED reference's
beliefs::bayes_adaptive::NestedBelief (src/beliefs/bayes-adaptive/NestedBelief.cpp) under seed "42" on a
tabular model (episodic tiger) and a factored one (episodic factored tiger, 3 irrelevant features, match-uniform
structure prior): the top filter (count blocks, structures, weights) and the bottom filters (domain states)
after initiate, and per script step the exact mt19937 words updateEstimation / resetDomainStateDistribution /
sample consumed and what they left behind.

Run from the repo root:  python oracle/gen_nested.py       (needs oracle/_ref/libfba_ref.so)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as GG  # noqa: E402
import pyref as R  # noqa: E402

CASES = {
    "tiger": (dict(domain="episodic-tiger", factored=False), 6, 40),
    "ftiger": (dict(domain="episodic-factored-tiger", size=3, factored=True, structure_prior="match-uniform"), 7, 33),
}
STEPS = 24


def main():
    out = {}
    for name, (cfg, n_top, n_bottom) in CASES.items():
        kw = dict(size=cfg.get("size", 0), factored=cfg["factored"], structure_prior=cfg.get("structure_prior", ""),
                  discount=GG.DISCOUNT, horizon=GG.HORIZON, seed="42")
        r = R.Ref(cfg["domain"], **kw)
        P = name + "/"
        for k, v in GG.model_desc(r, cfg).items():
            out[P + "model/" + k] = np.asarray(v)
        acts, obs, flags = r.env_script(STEPS, GG.HORIZON)
        out[P + "script/a"], out[P + "script/o"], out[P + "script/flags"] = acts, obs, flags
        r.reseed("47")
        r.composite_init(R.F_NESTED, n_top, n_bottom, 0.0)
        table = GG.StructTable()
        sid, _, counts = GG.dump_filter(r, R.F_NESTED, table)
        stride = counts.shape[1]
        w, tot = r.weights(R.F_NESTED)
        out[P + "init_struct_id"], out[P + "init_counts"], out[P + "init_w"] = sid, counts, w
        out[P + "init_total_weight"] = np.float64(tot)
        out[P + "init_states"] = r.nested_states(n_top, n_bottom)
        done = 0
        for t in range(STEPS):
            a, o, fl = int(acts[t]), int(obs[t]), int(flags[t])
            if fl & 2 and t > 0:
                r.mark()
                r.composite_reset(R.F_NESTED)
                out[P + "%d/reset_words" % t] = r.words_since_mark()
                out[P + "%d/reset_states" % t] = r.nested_states(n_top, n_bottom)
            if fl & 1:
                continue
            r.mark()
            r.composite_update(R.F_NESTED, a, o)
            out[P + "%d/words" % t] = r.words_since_mark()
            sid2, _, c = GG.dump_filter(r, R.F_NESTED, table, stride)
            assert np.array_equal(sid, sid2)
            w, tot = r.weights(R.F_NESTED)
            out[P + "%d/counts" % t], out[P + "%d/w" % t] = c, w
            out[P + "%d/total_weight" % t] = np.float64(tot)
            out[P + "%d/states" % t] = r.nested_states(n_top, n_bottom)
            r.mark()
            out[P + "%d/sample" % t] = np.array(r.nested_sample(), np.int32)
            out[P + "%d/sample_words" % t] = r.words_since_mark()
            done += 1
        out[P + "structs/t_par"], out[P + "structs/o_par"] = np.stack(table.t), np.stack(table.o)
        r.close()
        print("%s: %d x %d, stride %d, %d structures, %d updates" % (name, n_top, n_bottom, stride, len(table.t), done))
    np.savez_compressed(os.path.join(GG.OUT, "nested.npz"), **out)


if __name__ == "__main__":
    main()
