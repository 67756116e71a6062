#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generates tests/golden/composite.npz by running the UNMODIFIED reference's
composite structure beliefs under seed "42" on episodic-factored-tiger (3 irrelevant features,
match-uniform structure prior):

  cheat/…  beliefs::bayes_adaptive::prototypes::CheatingReinvigoration
           (src/beliefs/bayes-adaptive/prototypes/CheatingReinvigoration.cpp): both filters after initiate,
           then per script step the exact mt19937 words updateEstimation / resetDomainStateDistribution
           consumed and what they left behind (states, structures, count sums, weights, likelihood);
  inc/…    beliefs::bayes_adaptive::factored::StructureIncubatorSampling
           (src/beliefs/bayes-adaptive/factored/StructureIncubatorSampling.cpp): the three filters after
           initiate, per step words and results, and one step taken apart on a shadow belief with
           NON-uniform weights (promotion into the belief, leastLikely on distinct weights, replacement
           weights), which the class's own update never produces because it resamples every step.

Run from the repo root:  python oracle/gen_composite.py       (needs oracle/_ref/libfba_ref.so)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as GG  # noqa: E402
import pyref as R  # noqa: E402

N, SIZE, AMOUNT = 48, 3, 6
CHEAT_THRESHOLD = -1.5
INC_THRESHOLD = 0.05          # > 1 / N: the class's own updates promote nothing (as in the reference's runs)
STEPS = 40


def dump(r, out, key, filt, table, stride, weighted=False, full=False):
    sid, st, c = GG.dump_filter(r, filt, table, stride)
    out[key + "_struct_id"], out[key + "_state"] = sid, st
    out[key + "_count_sums"] = GG.count_sums(c)
    if full:
        out[key + "_counts"] = c
    if weighted:
        w, tot = r.weights(filt)
        out[key + "_w"], out[key + "_total_weight"] = w, np.float64(tot)


def main():
    cfg = dict(domain="episodic-factored-tiger", size=SIZE, factored=True, structure_prior="match-uniform")
    kw = dict(size=SIZE, factored=True, structure_prior="match-uniform", discount=GG.DISCOUNT, horizon=GG.HORIZON,
              seed="42")
    out = {}
    r = R.Ref(cfg["domain"], **kw)
    for k, v in GG.model_desc(r, cfg).items():
        out["model/" + k] = np.asarray(v)
    out["meta/discount"], out["meta/horizon"] = np.float64(GG.DISCOUNT), np.int32(GG.HORIZON)
    out["meta/N"], out["meta/amount"] = np.int32(N), np.int32(AMOUNT)
    out["meta/cheat_threshold"], out["meta/inc_threshold"] = np.float64(CHEAT_THRESHOLD), np.float64(INC_THRESHOLD)
    acts, obs, flags = r.env_script(STEPS, GG.HORIZON)
    out["script/a"], out["script/o"], out["script/flags"] = acts, obs, flags
    table = GG.StructTable()
    # the stride every filter uses: the fully connected structure's size
    FS, FO = len(r.feat_s), len(r.feat_o)

    # ---------------- CheatingReinvigoration ----------------
    r.reseed("45")
    r.composite_init(R.F_CHEAT, N, AMOUNT, CHEAT_THRESHOLD)
    # fully connected listen node bounds every structure factored tiger's mutate / prior can produce
    probe = [r.particle(R.F_CHEAT, i) for i in range(N)] + [r.particle(R.F_CHEAT_CORRECT, i) for i in range(N)]
    tp0, op0, _ = probe[0]
    op_full = op0.copy()
    op_full[2, 0] = (1 << FS) - 1
    stride = len(r.prior_model(tp0, op_full))
    out["meta/stride"] = np.int64(stride)
    dump(r, out, "cheat/init_b", R.F_CHEAT, table, stride, weighted=True, full=True)
    dump(r, out, "cheat/init_c", R.F_CHEAT_CORRECT, table, stride, full=True)
    for t in range(STEPS):
        a, o, fl = int(acts[t]), int(obs[t]), int(flags[t])
        if fl & 2 and t > 0:
            r.mark()
            r.composite_reset(R.F_CHEAT)
            out["cheat/%d/reset_words" % t] = r.words_since_mark()
            out["cheat/%d/reset_b_state" % t] = r.states(R.F_CHEAT)
            out["cheat/%d/reset_c_state" % t] = r.states(R.F_CHEAT_CORRECT)
        if fl & 1:
            continue
        r.mark()
        r.composite_update(R.F_CHEAT, a, o)
        out["cheat/%d/words" % t] = r.words_since_mark()
        lik = r.cheat_likelihood()
        out["cheat/%d/likelihood" % t] = np.float64(lik)
        dump(r, out, "cheat/%d/b" % t, R.F_CHEAT, table, stride, weighted=True)
        dump(r, out, "cheat/%d/c" % t, R.F_CHEAT_CORRECT, table, stride)
    dump(r, out, "cheat/final_b", R.F_CHEAT, table, stride, weighted=True, full=True)
    dump(r, out, "cheat/final_c", R.F_CHEAT_CORRECT, table, stride, full=True)
    n_cheat_updates = sum(1 for t in range(STEPS) if ("cheat/%d/likelihood" % t) in out
                          and out["cheat/%d/likelihood" % t] == 1.0)

    # ---------------- StructureIncubatorSampling ----------------
    r.reseed("46")
    r.mark()
    r.composite_init(R.F_INC, N, AMOUNT, INC_THRESHOLD)
    dump(r, out, "inc/init_b", R.F_INC, table, stride, full=True)
    dump(r, out, "inc/init_fc", R.F_INC_FC, table, stride, full=True)
    dump(r, out, "inc/init_s", R.F_INC_SHADOW, table, stride, weighted=True, full=True)
    # initiate's tail: the shadow belief bred from the two filters (StructureIncubatorSampling.cpp:74-80).
    # Re-run it alone on a second instance is impossible (private); instead record how many words the WHOLE
    # initiate took and let the tests replay the tail: the breeding is the last N breeds of the stream.
    out["inc/init_words"] = r.words_since_mark()
    done = 0
    for t in range(STEPS):
        a, o, fl = int(acts[t]), int(obs[t]), int(flags[t])
        if fl & 2 and t > 0:
            r.mark()
            r.composite_reset(R.F_INC)
            out["inc/%d/reset_words" % t] = r.words_since_mark()
            for tag, f in (("b", R.F_INC), ("fc", R.F_INC_FC), ("s", R.F_INC_SHADOW)):
                out["inc/%d/reset_%s_state" % (t, tag)] = r.states(f)
        if fl & 1:
            continue
        r.mark()
        r.composite_update(R.F_INC, a, o)
        out["inc/%d/words" % t] = r.words_since_mark()
        dump(r, out, "inc/%d/b" % t, R.F_INC, table, stride)
        dump(r, out, "inc/%d/fc" % t, R.F_INC_FC, table, stride)
        dump(r, out, "inc/%d/s" % t, R.F_INC_SHADOW, table, stride, weighted=True)
        done += 1
        if done >= 10:
            break
    out["inc/last_step"] = np.int32(t)
    # one more update taken apart, on distinct shadow weights: 5 heavy particles above the threshold
    w = 1.0 + 0.37 * ((np.arange(N) * 7) % 11)
    w[[3, 17, 18, 40, 41]] = 40.0
    r.incubator_set_shadow_weights(w)
    w0, tot0 = r.weights(R.F_INC_SHADOW)
    out["inc/parts/w_before"], out["inc/parts/total_before"] = w0, np.float64(tot0)
    dump(r, out, "inc/parts/before_b", R.F_INC, table, stride, full=True)
    dump(r, out, "inc/parts/before_fc", R.F_INC_FC, table, stride, full=True)
    dump(r, out, "inc/parts/before_s", R.F_INC_SHADOW, table, stride, weighted=True, full=True)
    t = int(out["inc/last_step"]) + 1
    while flags[t] & 1:
        t += 1
    a, o = int(acts[t]), int(obs[t])
    out["inc/parts/a"], out["inc/parts/o"] = np.int32(a), np.int32(o)
    for part in range(6):
        r.mark()
        lik = r.incubator_part(part, a, o)
        out["inc/parts/%d/words" % part] = r.words_since_mark()
        out["inc/parts/%d/likelihood" % part] = np.float64(lik)
        dump(r, out, "inc/parts/%d/b" % part, R.F_INC, table, stride, full=(part == 0))
        dump(r, out, "inc/parts/%d/fc" % part, R.F_INC_FC, table, stride)
        dump(r, out, "inc/parts/%d/s" % part, R.F_INC_SHADOW, table, stride, weighted=True, full=(part == 1))

    out["structs/t_par"], out["structs/o_par"] = np.stack(table.t), np.stack(table.o)
    np.savez_compressed(os.path.join(GG.OUT, "composite.npz"), **out)
    r.close()
    print("composite.npz: N=%d, stride=%d, %d structures, %d cheating updates that cheated, %d incubator updates"
          % (N, stride, len(table.t), n_cheat_updates, done))


if __name__ == "__main__":
    main()
