"""Statistical fixtures for the SAMPLED-Dirichlet mode (--dirichlet_sampling_method regular,
reference src/utils/random.cpp:189-242, 281-304; chosen in BAPOMDP.cpp:197-201 / FBAPOMDP.cpp:88-92).

That mode draws Gamma variates through the reference's ziggurat normal generator, so there is no
word-for-word replay contract for it: the CUDA path is checked STATISTICALLY against the
unmodified reference. This script runs the reference (oracle/_ref/libfba_ref.so, sampled mode) for
the first six belief updates of each golden script, REPS independent replicas of N particles, and
stores the replica means of the step likelihood and of the posterior feature marginals, plus their
standard errors, in tests/golden/sampled_stats.npz. Needs /root/reference (run in the build
container):   python oracle/gen_sampled_stats.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import pyref  # noqa: E402
import golden_util as G  # noqa: E402
from gen_golden import CONFIGS, DISCOUNT, HORIZON  # noqa: E402

NAMES = ["tiger", "ftiger", "sysadmin3", "ca", "gridworld3"]
N, REPS, STEPS = 3000, 6, 6


def marginals(state, w, fs):
    steps = np.concatenate([np.cumprod(fs[::-1])[::-1][1:], [1]])
    out = []
    for f in range(len(fs)):
        v = (state // steps[f]) % fs[f]
        out.append(np.bincount(v, weights=w, minlength=fs[f]))
    return np.concatenate(out)


def main():
    out = {}
    for name in NAMES:
        cfg, g = CONFIGS[name], G.load(name)
        script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)][:STEPS]
        fs = np.asarray(g.desc["feat_s"]).reshape(-1)
        lik = np.zeros((REPS, len(script)))
        marg = np.zeros((REPS, len(script), int(fs.sum())))
        for rep in range(REPS):
            r = pyref.Ref(cfg["domain"], size=cfg.get("size", 0), width=cfg.get("width", 0),
                          height=cfg.get("height", 0), factored=cfg["factored"],
                          structure_prior=cfg.get("structure_prior", ""), discount=DISCOUNT, horizon=HORIZON,
                          seed=str(1000 + rep), sampled=True)
            try:
                r.belief_init(pyref.F_IS, N)
                for t, (a, o) in enumerate(script):
                    lik[rep, t] = r.is_update(a, o)
                    w, _tot = r.is_weights()
                    marg[rep, t] = marginals(r.states(pyref.F_IS), w / w.sum(), fs)
                    r.is_resample()
            finally:
                r.close()
        out[name + "/script"] = np.array(script, np.int32)
        out[name + "/lik_mean"], out[name + "/lik_se"] = lik.mean(0), lik.std(0, ddof=1) / np.sqrt(REPS)
        out[name + "/marg_mean"], out[name + "/marg_se"] = marg.mean(0), marg.std(0, ddof=1) / np.sqrt(REPS)
        print(name, "lik", np.round(lik.mean(0), 4), "se", np.round(out[name + "/lik_se"], 4))
    out["meta/N"], out["meta/reps"] = np.int32(N), np.int32(REPS)
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "sampled_stats.npz"), **out)


if __name__ == "__main__":
    main()
