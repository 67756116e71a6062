"""Golden vectors for BABNModel::LogBDScore (BABNModel.cpp:451-478, DBNNode.cpp:82-117): particles of an
importance-sampling belief after a few updates, scored by the UNMODIFIED reference against the prior
particle of a second (fresh) belief with the same structure. -> tests/golden/bd_score.npz
Needs /root/reference (run in the build container):   python oracle/gen_bd_score.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import pyref  # noqa: E402
import golden_util as G  # noqa: E402
from gen_golden import CONFIGS, DISCOUNT, HORIZON  # noqa: E402

out = {}
for name in ("ftiger", "sysadmin3", "sysadmin", "gridworld3_fba"):
    cfg, g = CONFIGS[name], G.load(name)
    r = pyref.Ref(cfg["domain"], size=cfg.get("size", 0), width=cfg.get("width", 0), height=cfg.get("height", 0),
                  factored=True, structure_prior=cfg.get("structure_prior", ""), discount=DISCOUNT, horizon=HORIZON,
                  seed="17")
    n = 12
    r.belief_init(pyref.F_IS, n)
    r.belief_init(pyref.F_RS, 1)         # the prior particle (no structure prior: one structure for all)
    script = [(int(a), int(o)) for a, o, f in zip(g.a, g.o, g.flags) if not (f & 1)][:8]
    for a, o in script:
        r.update_estimation(pyref.F_IS, a, o)
    tp0, op0, prior = r.particle(pyref.F_RS, 0)
    counts, scores = [], []
    for i in range(n):
        tp, op, c = r.particle(pyref.F_IS, i)
        assert np.array_equal(tp, tp0) and np.array_equal(op, op0)
        counts.append(c)
        scores.append(r.log_bd_score(pyref.F_IS, i, pyref.F_RS, 0))
    r.close()
    out[name + "/t_par"], out[name + "/o_par"] = tp0, op0
    out[name + "/prior"], out[name + "/counts"] = prior, np.stack(counts)
    out[name + "/score"] = np.array(scores, np.float64)
    print(name, np.round(scores[:4], 6))
np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "bd_score.npz"), **out)
