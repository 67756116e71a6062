#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generates tests/golden/mutate.npz — chains of FBAPOMDP::mutate (the domain priors' mutate,
src/bayes-adaptive/models/factored/FBAPOMDP.cpp:57-61) run by the UNMODIFIED reference under seed "42" on the four
domains that have one: every step's input structure, the exact mt19937 words it consumed and the structure it
returned. Pins the gridworld mutate (GridWorldBAPriors.cpp:200-225), which the reference's reinvigoration cannot reach
(sampleFullyConnectedState is "nyi" there), and the others once more, directly.

Run from the repo root:  python oracle/gen_mutate.py       (needs oracle/_ref/libfba_ref.so)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as GG  # noqa: E402
import pyref as R  # noqa: E402

CASES = {
    # name: (reference flags, mutate kind of include/fba_pomdp_b200.h)
    "gridworld": (dict(domain="gridworld", size=3, factored=True), 3),
    "ftiger": (dict(domain="episodic-factored-tiger", size=4, factored=True), 0),
    "ca": (dict(domain="centered-collision-avoidance", size=2, width=5, height=5, factored=True), 1),
    "sysadmin": (dict(domain="linear-sysadmin", size=5, factored=True), 2),
}
STEPS = 48


def main():
    out = {}
    for name, (cfg, kind) in CASES.items():
        r = R.Ref(cfg["domain"], size=cfg.get("size", 0), width=cfg.get("width", 0), height=cfg.get("height", 0),
                  factored=True, discount=GG.DISCOUNT, horizon=GG.HORIZON, seed="42")
        P = name + "/"
        for k, v in GG.model_desc(r, cfg).items():
            out[P + "model/" + k] = np.asarray(v)
        out[P + "kind"] = np.int32(kind)
        r.belief_init(R.F_IS, 1)
        tp, op, _ = r.particle(R.F_IS, 0)                       # the prior's own structure as the chain's start
        tp, op = tp.reshape(-1), op.reshape(-1)
        t_in, o_in, t_out, o_out, words, n_words = [], [], [], [], [], []
        r.reseed("49")
        for _ in range(STEPS):
            r.mark()
            tp2, op2 = r.mutate(tp, op)
            w = r.words_since_mark()
            t_in.append(tp), o_in.append(op), t_out.append(tp2), o_out.append(op2)
            words.append(w), n_words.append(len(w))
            tp, op = tp2, op2
        out[P + "t_in"], out[P + "o_in"] = np.stack(t_in), np.stack(o_in)
        out[P + "t_out"], out[P + "o_out"] = np.stack(t_out), np.stack(o_out)
        out[P + "words"], out[P + "n_words"] = np.concatenate(words), np.array(n_words, np.int64)
        r.close()
        print("%s: %d mutations, %d words, %d distinct structures" % (name, STEPS, sum(n_words),
              len({(a.tobytes(), b.tobytes()) for a, b in zip(t_out, o_out)})))
    np.savez_compressed(os.path.join(GG.OUT, "mutate.npz"), **out)


if __name__ == "__main__":
    main()
