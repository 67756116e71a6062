"""GPU: POMCP with the search tree on the device (fba_tree_*). The simulations of a
wave share the tree through atomics, so results are not replayable word for word; the checks are
(a) exact where the answer is known in closed form (one-step action values of the tiger problem,
visit counts, argument errors) and (b) statistical: root action values agree between a sequential
search (wave = 1: the reference's algorithm) and a wide one, between dense and base+delta storage,
and the chosen action is the informed one. Episode-level parity against the reference's own planner
is in test_dropin_adapter.py."""
import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import fba_pomdp_b200 as fba
    c = fba.Context(0)
    yield c
    c.close()


def _tiger(ctx, n, states, delta=0, weighted=True):
    import fba_pomdp_b200 as fba
    g = G.load("tiger")
    desc = dict(g.desc)
    desc["delta_capacity"] = delta
    sim = fba.BAPOMDP(ctx, desc, g.t_par, g.o_par)
    b = fba.BAImportanceSampling(n) if weighted else fba.BARejectionSampling(n)
    b.initiate(sim, proto_struct_id=[0], proto_counts=g["is/init_counts"][:1], particle_proto=None,
               state=np.asarray(states, np.int32))
    return g, sim, b


def test_one_step_values_are_exact(ctx):
    """depth 1: every simulation is one step from the root, so Q(listen) = -1 exactly and
    Q(open a) = +10 when the tiger is behind the other door in every particle (Tiger rewards,
    TigerBAExtension.cpp:21-44)."""
    import fba_pomdp_b200 as fba
    n = 512
    g, sim, b = _tiger(ctx, n, np.zeros(n))          # state 0 in every particle
    tree = fba.SearchTree(sim, 3000, 4)
    a, q, visits = tree.selectAction(b, 3000, 1, 100.0, 0.95, 64, fba.Rng.philox(1))
    assert visits.sum() == 3000 and visits.min() >= 1
    assert q[2] == -1.0 and q[0] == 10.0 and q[1] == -100.0 and a == 0
    # UCB at work: the bad door is tried only while unvisited (first wave), and the visits of the two
    # reasonable actions balance their upper confidence bounds q + u sqrt(log(N + 1) / n)
    assert visits[1] <= 64 and visits[1] < visits[2] < visits[0]
    ucb = q + 100.0 * np.sqrt(np.log1p(3000.0) / visits)
    assert abs(ucb[0] - ucb[2]) < 2.0, ucb
    # depth 0: nothing to simulate, every value is 0
    a0, q0, v0 = tree.selectAction(b, 100, 0, 100.0, 0.95, 64, fba.Rng.philox(2))
    assert v0.sum() == 0 and np.all(q0 == 0) and 0 <= a0 < 3
    tree.free()
    b.free()
    sim.close()


def test_wave_width_trades_search_depth_for_latency(ctx):
    """Tiger with an uninformed belief: opening a door is worth (10 - 100) / 2 = -45 whatever the
    search does afterwards (the episode ends) — every wave width must find that, within 6 standard
    errors of a -100/+10 coin (sd 55) at the visits the action got. Listening is worth -1 plus the
    discounted value of what the tree policy does next: a sequential search (wave = 1, the
    reference's algorithm) learns to listen again and open the right door, while the simulations of
    one wide wave cannot learn from each other and mostly fall through to random rollouts (which
    open doors blindly): Q(listen) is clearly better for wave 1 than for one 4096-wide wave, with
    256 in between — the documented price of width."""
    import fba_pomdp_b200 as fba
    n = 2048
    rs = np.random.RandomState(0)
    g, sim, b = _tiger(ctx, n, rs.randint(0, 2, n))
    tree = fba.SearchTree(sim, 8192, 8)
    out, vis = {}, {}
    for wave in (1, 256, 4096):
        qs, vs = [], []
        for rep in range(4):
            sims = 1024 if wave == 1 else 4096
            a, q, visits = tree.selectAction(b, sims, 4, 50.0, 0.95, wave, fba.Rng.philox(100 * wave + rep))
            assert visits.sum() == sims and 0 <= a < 3
            qs.append(q * visits)
            vs.append(visits)
        vis[wave] = np.sum(vs, axis=0)
        out[wave] = np.sum(qs, axis=0) / vis[wave]
    for wave in (1, 256, 4096):
        se = 55.0 / np.sqrt(vis[wave][:2])
        assert np.all(np.abs(out[wave][:2] + 45.0) < 6.0 * se), (wave, out, vis)
    assert out[1][2] > out[256][2] - 2.0 and out[256][2] > out[4096][2] - 2.0 and out[1][2] > out[4096][2] + 5.0, out
    tree.free()
    b.free()
    sim.close()


def _listen_then_random_value(depth, discount=0.95):
    """Q(listen) when every later action is uniformly random, tiger position known to the model but not
    used: V_0 = 0, V_d = (-100 + 10 + (-1 + discount V_{d-1})) / 3 (opening ends the episode,
    TigerBAExtension.cpp:21-44), Q = -1 + discount V_{depth-1}."""
    v = 0.0
    for _ in range(depth - 1):
        v = (-100.0 + 10.0 + (-1.0 + discount * v)) / 3.0
    return -1.0 + discount * v


def test_flat_and_delta_beliefs(ctx):
    """A flat (rejection-sampling) belief and a base+delta stored belief search like the dense
    weighted one: same informed choice, exact terminal values, and a value of listening inside the
    bounds the model gives. Bounds on Q(listen), depth 3, discount 0.95:
      upper: -1 + 0.95 * 10 = 8.5 (listen once, then the right door);
      lower: the tree policy below the root is UCB over tried actions after each action was tried once,
             so in expectation it is no worse than the uniformly random policy, whose value is
             _listen_then_random_value(3) = -38.9.
    The simulations of one wave are NOT independent samples (they share statistics while in flight), so the unit
    of the statistical check is the SEARCH: 64 seeds per storage kind, the mean of their Q(listen) must exceed
    -38.9 minus 4 standard errors of that mean (estimated from the 64 values) minus 2 (measured: mean -39.0,
    per-search sd 3.7, profiles/r2w_tree_seeds.txt; round 1's herding bug gave -78). No single search is judged
    on its own — 8 seeds and a bound that assumed independent simulations failed once in twenty suite runs."""
    import fba_pomdp_b200 as fba
    n = 1024
    states = np.ones(n)          # tiger right everywhere: open-right (action 1) pays +10
    q_random = _listen_then_random_value(3)
    assert abs(q_random + 38.94) < 0.01
    for kw in (dict(), dict(weighted=False), dict(delta=64)):
        g, sim, b = _tiger(ctx, n, states, **kw)
        tree = fba.SearchTree(sim, 4096, 6)
        q_listen = []
        for seed in range(64):
            a, q, visits = tree.selectAction(b, 4096, 3, 30.0, 0.95, 512, fba.Rng.philox(7 + seed))
            assert a == 1 and visits.sum() == 4096 and visits.min() >= 1
            # tiger is episodic here (opening ends the episode): exact terminal values
            assert q[1] == 10.0 and q[0] == -100.0
            assert -96.0 <= q[2] <= 8.5
            q_listen.append(q[2])
        q_listen = np.array(q_listen)
        se = q_listen.std(ddof=1) / np.sqrt(len(q_listen))
        assert q_listen.mean() > q_random - 4.0 * se - 2.0, (kw, q_listen.mean(), se, q_listen.min())
        tree.free()
        b.free()
        sim.close()


def test_tree_argument_checks(ctx):
    import fba_pomdp_b200 as fba
    g, sim, b = _tiger(ctx, 64, np.zeros(64))
    tree = fba.SearchTree(sim, 128, 4)
    with pytest.raises(fba.FbaError):
        tree.selectAction(b, 129, 2, 1.0, 0.95, 16, fba.Rng.philox(1))       # more than max_simulations
    with pytest.raises(fba.FbaError):
        tree.selectAction(b, 64, 5, 1.0, 0.95, 16, fba.Rng.philox(1))        # deeper than max_depth
    with pytest.raises(fba.FbaError):
        tree.selectAction(b, 64, 2, 1.0, 0.95, 16, fba.Rng.replay(np.zeros(8, np.uint32)))
    with pytest.raises(fba.FbaError):
        fba.SearchTree(sim, 0, 4)
    tree.free()
    b.free()
    sim.close()
