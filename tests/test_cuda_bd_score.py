"""GPU: BABNModel::LogBDScore on the device (fba_belief_log_bd_score) against the unmodified reference's
values (tests/golden/bd_score.npz, oracle/gen_bd_score.py) and the bit-exact CPU oracle. The device sums
in another order and uses CUDA's lgamma: tolerance 1e-10 relative (+1e-10 absolute)."""
import os

import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ftiger", "sysadmin3", "sysadmin", "gridworld3_fba"])
def test_log_bd_score_vs_reference(name):
    import fba_pomdp_b200 as fba
    z = np.load(os.path.join(G.GOLDEN_DIR, "bd_score.npz"))
    g = G.load(name)
    ctx = fba.Context(0)
    sim = fba.BAPOMDP(ctx, g.desc, z[name + "/t_par"][None], z[name + "/o_par"][None])
    counts, prior, want = z[name + "/counts"], z[name + "/prior"], z[name + "/score"]
    n = len(counts)
    b = fba.BAImportanceSampling(n)
    b.initiate(sim, struct_id=np.zeros(n, np.int32), counts=counts, state=np.zeros(n, np.int32))
    p1 = fba.BAImportanceSampling(1)
    p1.initiate(sim, struct_id=[0], counts=prior[None, :], state=[0])
    pn = fba.BAImportanceSampling(n)
    pn.initiate(sim, struct_id=np.zeros(n, np.int32), counts=np.tile(prior, (n, 1)), state=np.zeros(n, np.int32))
    for pr in (p1, pn):
        got = fba.log_bd_score(b, pr)
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)
    # a particle scored against itself: exactly 0
    np.testing.assert_array_equal(fba.log_bd_score(b, b), np.zeros(n))
    wrong = fba.BAImportanceSampling(3)
    wrong.initiate(sim, struct_id=np.zeros(3, np.int32), counts=counts[:3], state=np.zeros(3, np.int32))
    with pytest.raises(fba.FbaError):
        fba.log_bd_score(b, wrong)
    for x in (b, p1, pn, wrong):
        x.free()
    sim.close()
    ctx.close()
