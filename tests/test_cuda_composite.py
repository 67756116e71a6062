"""GPU: the composite structure beliefs (fba-pomdp_b200/structure_beliefs.py over the C ABI) in REPLAY mode
against the UNMODIFIED reference's own CheatingReinvigoration and StructureIncubatorSampling through
tests/golden/composite.npz (oracle/gen_composite.py): fed the exact mt19937 words each call consumed, every
updateEstimation / resetDomainStateDistribution leaves the reference's filters behind bit for bit — domain
states, structures, counts, weights, _total_weight, accumulated likelihood — and consumes every word."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "composite.npz")
MUT_FACTORED_TIGER = 0


@pytest.fixture(scope="module")
def env():
    import fba_pomdp_b200 as fba
    g = np.load(GOLDEN)
    desc = {k[len("model/"):]: g[k] for k in g.files if k.startswith("model/")}
    ctx = fba.Context(0)
    sim = fba.BAPOMDP(ctx, desc, g["structs/t_par"], g["structs/o_par"], max_structures=len(g["structs/t_par"]) + 512)
    yield fba, g, sim
    sim.close()
    ctx.close()


def particles(g, prefix):
    return dict(struct_id=g[prefix + "_struct_id"], counts=g[prefix + "_counts"], state=g[prefix + "_state"])


def assert_matches(b, g, prefix, weighted=False):
    d = b.download()
    np.testing.assert_array_equal(d["state"], g[prefix + "_state"], err_msg=prefix)
    np.testing.assert_array_equal(d["struct_id"], g[prefix + "_struct_id"], err_msg=prefix)
    np.testing.assert_array_equal(d["counts"].astype(np.float64).sum(1), g[prefix + "_count_sums"], err_msg=prefix)
    if (prefix + "_counts") in g.files:
        np.testing.assert_array_equal(d["counts"], g[prefix + "_counts"], err_msg=prefix)
    if weighted:
        np.testing.assert_array_equal(d["w"], g[prefix + "_w"], err_msg=prefix)
        assert d["total_weight"] == float(g[prefix + "_total_weight"]), prefix


def set_weights(b, w, total):
    from fba_pomdp_b200.capi import ptr
    w = np.ascontiguousarray(w, np.float64)
    rc = b.L.fba_belief_upload(b.h, 0, len(w), None, None, None, ptr(w))
    assert rc == 0
    assert b.download(counts=False)["total_weight"] == total   # upload re-accumulates in order, as the harness did


def test_cheating_reinvigoration_replay(env):
    fba, g, sim = env
    from fba_pomdp_b200.structure_beliefs import CheatingReinvigoration
    N, stride = int(g["meta/N"]), int(g["meta/stride"])
    b = CheatingReinvigoration(N, int(g["meta/amount"]), float(g["meta/cheat_threshold"]))
    b.initiate(sim, belief=particles(g, "cheat/init_b"), correct_structured=particles(g, "cheat/init_c"), stride=stride)
    a_, o_, fl_ = g["script/a"], g["script/o"], g["script/flags"]
    updates = 0
    for t in range(len(a_)):
        if fl_[t] & 2 and t > 0:
            rng = fba.Rng.replay(g["cheat/%d/reset_words" % t])
            b.resetDomainStateDistribution(rng)
            assert rng.exhausted
            np.testing.assert_array_equal(b._belief.download(counts=False)["state"], g["cheat/%d/reset_b_state" % t])
            np.testing.assert_array_equal(b._correct_structured_belief.download(counts=False)["state"],
                                          g["cheat/%d/reset_c_state" % t])
        if fl_[t] & 1:
            continue
        rng = fba.Rng.replay(g["cheat/%d/words" % t])
        b.updateEstimation(int(a_[t]), int(o_[t]), rng)
        assert rng.exhausted, t
        assert b._likelihood == float(g["cheat/%d/likelihood" % t])
        assert_matches(b._belief, g, "cheat/%d/b" % t, weighted=True)
        assert_matches(b._correct_structured_belief, g, "cheat/%d/c" % t)
        updates += 1
    assert_matches(b._belief, g, "cheat/final_b", weighted=True)
    assert_matches(b._correct_structured_belief, g, "cheat/final_c")
    assert updates >= 12 and b.cheats >= 3
    b.free()


def test_structure_incubator_replay(env):
    fba, g, sim = env
    from fba_pomdp_b200.structure_beliefs import StructureIncubatorSampling
    N, stride = int(g["meta/N"]), int(g["meta/stride"])
    inc = StructureIncubatorSampling(N, int(g["meta/amount"]), float(g["meta/inc_threshold"]), MUT_FACTORED_TIGER)
    inc.initiate(sim, belief=particles(g, "inc/init_b"), fully_connected=particles(g, "inc/init_fc"), stride=stride,
                 shadow=particles(g, "inc/init_s"))
    set_weights(inc._shadow_belief, g["inc/init_s_w"], float(g["inc/init_s_total_weight"]))
    a_, o_, fl_ = g["script/a"], g["script/o"], g["script/flags"]
    done = 0
    for t in range(int(g["inc/last_step"]) + 1):
        if fl_[t] & 2 and t > 0:
            rng = fba.Rng.replay(g["inc/%d/reset_words" % t])
            inc.resetDomainStateDistribution(rng)
            assert rng.exhausted
            for tag, b in (("b", inc._belief), ("fc", inc._fully_connected_belief), ("s", inc._shadow_belief)):
                np.testing.assert_array_equal(b.download(counts=False)["state"], g["inc/%d/reset_%s_state" % (t, tag)])
        if fl_[t] & 1:
            continue
        rng = fba.Rng.replay(g["inc/%d/words" % t])
        inc.updateEstimation(int(a_[t]), int(o_[t]), rng)
        assert rng.exhausted, t
        assert_matches(inc._belief, g, "inc/%d/b" % t)
        assert_matches(inc._fully_connected_belief, g, "inc/%d/fc" % t)
        assert_matches(inc._shadow_belief, g, "inc/%d/s" % t, weighted=True)
        done += 1
    assert done == 10
    inc.free()


def test_incubator_parts_on_distinct_shadow_weights(env):
    """promotion into the belief, leastLikely and WeightedFilter::replace weights on NON-uniform shadow weights"""
    fba, g, sim = env
    from fba_pomdp_b200.structure_beliefs import StructureIncubatorSampling
    N, stride = int(g["meta/N"]), int(g["meta/stride"])
    inc = StructureIncubatorSampling(N, int(g["meta/amount"]), float(g["meta/inc_threshold"]), MUT_FACTORED_TIGER)
    inc.initiate(sim, belief=particles(g, "inc/parts/before_b"), fully_connected=particles(g, "inc/parts/before_fc"),
                 stride=stride, shadow=particles(g, "inc/parts/before_s"))
    set_weights(inc._shadow_belief, g["inc/parts/before_s_w"], float(g["inc/parts/before_s_total_weight"]))
    a, o = int(g["inc/parts/a"]), int(g["inc/parts/o"])
    steps = [inc.reinvigorateBelief, inc.reinvigorateShadowBelief,
             lambda r: inc._belief.updateEstimation(a, o, r), lambda r: inc._fully_connected_belief.updateEstimation(a, o, r),
             lambda r: inc._shadow_belief.update(a, o, r), lambda r: inc._shadow_belief.resample(r)]
    for part, fn in enumerate(steps):
        rng = fba.Rng.replay(g["inc/parts/%d/words" % part])
        res = fn(rng)
        assert rng.exhausted, part
        if part == 0:
            assert res == 5
        if part == 4:
            assert res == float(g["inc/parts/4/likelihood"])
        assert_matches(inc._belief, g, "inc/parts/%d/b" % part)
        assert_matches(inc._fully_connected_belief, g, "inc/parts/%d/fc" % part)
        assert_matches(inc._shadow_belief, g, "inc/parts/%d/s" % part, weighted=True)
    inc.free()


def test_incubator_initiate_breeds_the_shadow_belief(env):
    """StructureIncubatorSampling::initiate's tail (:74-80): N breeds into the shadow belief with weights
    1 / N — here in PHILOX mode: every shadow particle has a structure one mutation away from a belief
    particle's, the domain state of a belief particle, and counts that marginalise a fully connected
    particle's (equal row sums per node)."""
    fba, g, sim = env
    from fba_pomdp_b200.structure_beliefs import StructureIncubatorSampling
    N, stride = int(g["meta/N"]), int(g["meta/stride"])
    inc = StructureIncubatorSampling(N, int(g["meta/amount"]), float(g["meta/inc_threshold"]), MUT_FACTORED_TIGER)
    inc.initiate(sim, belief=particles(g, "inc/init_b"), fully_connected=particles(g, "inc/init_fc"), stride=stride,
                 rng=fba.Rng.philox(11))
    d = inc._shadow_belief.download()
    np.testing.assert_array_equal(d["w"], np.full(N, 1.0 / N))
    fc_sums = set(np.round(g["inc/init_fc_counts"].astype(np.float64).sum(1), 3).tolist())
    assert set(np.round(d["counts"].astype(np.float64).sum(1), 3).tolist()) <= fc_sums   # marginalising keeps the mass
    assert set(d["state"].tolist()) <= set(g["inc/init_b_state"].tolist())
    belief_structs = {(sim.structure(int(i))[0].tobytes(), sim.structure(int(i))[1].tobytes())
                      for i in np.unique(g["inc/init_b_struct_id"])}
    for i in np.unique(d["struct_id"]):
        t, o = sim.structure(int(i))
        one_flip = False
        for bt, bo in belief_structs:
            diff = np.frombuffer(bo, np.uint32) ^ o.reshape(-1)
            if bt == t.tobytes() and np.count_nonzero(diff) == 1 and bin(int(diff[diff != 0][0])).count("1") == 1:
                one_flip = True
        assert one_flip
    inc.free()


def test_replace_from_and_argument_checks(env):
    fba, g, sim = env
    from fba_pomdp_b200.structure_beliefs import replace_from
    N, stride = int(g["meta/N"]), int(g["meta/stride"])
    src = fba.BARejectionSampling(N)
    src.initiate(sim, **particles(g, "cheat/init_c"), stride=stride)
    dst = fba.BAImportanceSampling(N)
    dst.initiate(sim, **particles(g, "cheat/init_b"), stride=stride)
    before = dst.download()
    s = src.download()
    replace_from(dst, [5, 7, 5], src, [1, 2, 3])     # slot 5 is written twice: the later one stays
    d = dst.download()
    want = before["counts"].copy()
    want[7], want[5] = s["counts"][2], s["counts"][3]
    np.testing.assert_array_equal(d["counts"], want)
    assert d["state"][5] == s["state"][3] and d["struct_id"][7] == s["struct_id"][2]
    # WeightedFilter::replace weights, slot after slot
    w, tot = before["w"].copy(), before["total_weight"]
    for slot in (5, 7, 5):
        nw = tot / N
        tot = tot + (nw - w[slot])
        w[slot] = nw
    np.testing.assert_array_equal(d["w"], w)
    assert d["total_weight"] == tot
    with pytest.raises(fba.FbaError):
        replace_from(dst, [N], src, [0])
    with pytest.raises(fba.FbaError):
        replace_from(dst, [0], dst, [1])
    src.free()
    dst.free()


# ---- NestedBelief -----------------------------------------------------------------------------------------

NESTED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nested.npz")


def nested_env(name):
    import fba_pomdp_b200 as fba
    import pycomposite as PC
    import pyoracle as O
    g = np.load(NESTED)
    P = name + "/"
    desc = {k[len(P + "model/"):]: g[k] for k in g.files if k.startswith(P + "model/")}
    m = O.Model(desc)
    st = O.Structs(m, g[P + "structs/t_par"], g[P + "structs/o_par"])
    counts = g[P + "init_counts"]
    top = O.Belief(counts.shape[0], counts.shape[1], True)
    top.counts[:] = counts
    top.struct_id[:] = g[P + "init_struct_id"]
    top.w[:] = g[P + "init_w"]
    top.total_weight = float(g[P + "init_total_weight"])
    oracle = PC.Nested(m, st, top, g[P + "init_states"])
    ctx = fba.Context(0)
    sim = fba.BAPOMDP(ctx, desc, g[P + "structs/t_par"], g[P + "structs/o_par"])
    return fba, O, g, P, oracle, ctx, sim


@pytest.mark.parametrize("name", ["tiger", "ftiger"])
def test_nested_belief_replay(name):
    """fba_nested_* in REPLAY mode against the oracle's NestedBelief (pinned to the reference's class bit for bit by
    tests/test_oracle_composite.py). The reference consumes ONE stream particle after particle, a data-dependent
    amount each; the kernel gives top particle i the i-th equal slice of the stream, and the oracle is fed the same
    slices: bottom filters, counts, attempts, top weights and _total_weight identical. Resets and sample() consume
    the reference's own words and reproduce the reference's results."""
    from fba_pomdp_b200.structure_beliefs import NestedBelief
    fba, O, g, P, oracle, ctx, sim = nested_env(name)
    n_top, n_bottom = g[P + "init_states"].shape
    nb = NestedBelief(n_top, n_bottom)
    nb.initiate(sim, struct_id=g[P + "init_struct_id"], counts=g[P + "init_counts"], states=g[P + "init_states"])
    a_, o_, fl_ = g[P + "script/a"], g[P + "script/o"], g[P + "script/flags"]
    rs = np.random.RandomState(5)
    done = 0
    for t in range(len(a_)):
        if fl_[t] & 2 and t > 0:
            rng = fba.Rng.replay(g[P + "%d/reset_words" % t])
            nb.resetDomainStateDistribution(rng)
            assert rng.exhausted
            np.testing.assert_array_equal(nb.download()["states"], g[P + "%d/reset_states" % t])
            oracle.states[:] = g[P + "%d/reset_states" % t]
        if fl_[t] & 1:
            continue
        per = 40000
        words = rs.randint(0, 1 << 32, size=per * n_top, dtype=np.uint64).astype(np.uint32)
        rng = fba.Rng.replay(words)
        nb.updateEstimation(int(a_[t]), int(o_[t]), rng)
        oracle.update(int(a_[t]), int(o_[t]), O.Rng(words), slice_words=per)
        d = nb.download()
        np.testing.assert_array_equal(nb.attempts, oracle.attempts)
        np.testing.assert_array_equal(d["states"], oracle.states)
        np.testing.assert_array_equal(d["counts"], oracle.top.counts)
        np.testing.assert_array_equal(d["w"], oracle.top.w)
        assert d["total_weight"] == oracle.top.total_weight
        # sample(): the reference's words on the same weights give the oracle's draw
        sw = g[P + "%d/sample_words" % t]
        rng = fba.Rng.replay(sw)
        assert nb.sample(rng) == oracle.sample(O.Rng(sw))
        assert rng.exhausted
        done += 1
    assert done >= 8
    nb.free()
    sim.close()
    ctx.close()


def test_nested_belief_philox_matches_the_oracle_statistically():
    """PHILOX mode, many top particles from one prior: acceptance counts per top particle and the bottom filters'
    state histogram after an update agree with the oracle's within 5 standard errors; count blocks gain exactly
    n_bottom x (1 / n_bottom) per node; weights are normalised."""
    from fba_pomdp_b200.structure_beliefs import NestedBelief
    fba, O, g, P, oracle, ctx, sim = nested_env("tiger")
    n_top, n_bottom = 512, 64
    counts = np.repeat(g[P + "init_counts"][:1], n_top, 0)
    states = np.random.RandomState(1).randint(0, 2, size=(n_top, n_bottom)).astype(np.int32)
    a, o = 2, 0       # listen, hear left
    # the default PHILOX kernel (a warp per top particle, 32 attempts per round) and the attempt-by-attempt loop
    # (option "nested_exact") against each other: same acceptance statistics
    per_kernel = []
    for exact in (1, 0):
        ctx.set_option("nested_exact", exact)
        nb = NestedBelief(n_top, n_bottom)
        nb.initiate(sim, struct_id=np.zeros(n_top, np.int32), counts=counts, states=states)
        nb.updateEstimation(a, o, fba.Rng.philox(77))
        per_kernel.append((nb.attempts.copy(), nb.download()["states"].sum(1)))
        if exact:
            nb.free()
    for got, want in zip(per_kernel[0], per_kernel[1]):
        se = np.sqrt(got.var(ddof=1) / n_top + want.var(ddof=1) / n_top) + 1e-12
        assert abs(got.mean() - want.mean()) <= 5 * se, (got.mean(), want.mean(), se)
    d = nb.download()
    # oracle on the same inputs, its own words
    top = O.Belief(n_top, counts.shape[1], True)
    top.counts[:] = counts
    top.total_weight = O.sequential_uniform_total(n_top)
    import pycomposite as PC
    ob = PC.Nested(oracle.m, oracle.st, top, states)
    words = np.random.RandomState(2).randint(0, 1 << 32, size=4_000_000, dtype=np.uint64).astype(np.uint32)
    ob.update(a, o, O.Rng(words))
    for got, want in ((nb.attempts, ob.attempts), (d["states"].sum(1), ob.states.sum(1))):
        se = np.sqrt(got.var(ddof=1) / n_top + want.var(ddof=1) / n_top) + 1e-12
        assert abs(got.mean() - want.mean()) <= 5 * se, (got.mean(), want.mean(), se)
    gain = d["counts"].astype(np.float64).sum(1) - counts.astype(np.float64).sum(1)
    np.testing.assert_allclose(gain, 2.0, rtol=1e-4)          # one transition + one observation cell, 64 x 1/64 each
    assert abs(d["w"].sum() - 1.0) < 1e-12 and d["total_weight"] == pytest.approx(1.0, abs=1e-12)
    with pytest.raises(fba.FbaError):
        nb.updateEstimation(a, o, fba.Rng.philox(78), max_attempts=3)
    nb.free()
    sim.close()
    ctx.close()


MUTATE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mutate.npz")


@pytest.mark.parametrize("name", ["gridworld", "ftiger", "ca", "sysadmin"])
def test_domain_mutate_replay(name):
    """The C ABI's mutate (inside fba_belief_breed_into / _reinvigorate) against the reference's FBAPOMDP::mutate on the
    four domains that have one, the gridworld one included (the reference's reinvigoration cannot reach it): one
    breed per recorded mutation — a one-particle structure donor holding the input structure, a one-particle fully
    connected counts donor — fed [2 donor-draw words | the reference's mutate words]: the bred particle has the
    reference's output structure and every word is consumed."""
    import fba_pomdp_b200 as fba
    from fba_pomdp_b200.capi import ptr
    import ctypes as C
    g = np.load(MUTATE)
    P = name + "/"
    desc = {k[len(P + "model/"):]: g[k] for k in g.files if k.startswith(P + "model/")}
    A, FS, FO = int(desc["A"]), len(desc["feat_s"]), len(desc["feat_o"])
    full = (1 << FS) - 1
    t_all = np.concatenate([np.full((1, A * FS), full, np.uint32), g[P + "t_in"], g[P + "t_out"]])
    o_all = np.concatenate([np.full((1, A * FO), full, np.uint32), g[P + "o_in"], g[P + "o_out"]])
    ctx = fba.Context(0)
    sim = fba.BAPOMDP(ctx, desc, t_all, o_all, max_structures=len(t_all) + 8)
    key = {}
    for i in range(sim.num_structures):
        t, o = sim.structure(i)
        key[(t.reshape(-1).tobytes(), o.reshape(-1).tobytes())] = i
    stride = sim.max_structure_size()
    fc = fba.BARejectionSampling(1)
    fc.initiate(sim, struct_id=np.array([key[(t_all[0].tobytes(), o_all[0].tobytes())]], np.int32),
                counts=np.ones((1, stride), np.float32), state=np.zeros(1, np.int32), stride=stride)
    words, pos = g[P + "words"], 0
    for k in range(len(g[P + "t_in"])):
        nw = int(g[P + "n_words"][k])
        donor = fba.BARejectionSampling(1)
        donor.initiate(sim, struct_id=np.array([key[(g[P + "t_in"][k].tobytes(), g[P + "o_in"][k].tobytes())]], np.int32),
                       counts=np.ones((1, stride), np.float32), state=np.array([1], np.int32), stride=stride)
        dst = fba.BARejectionSampling(1)
        dst.initiate(sim, struct_id=np.zeros(1, np.int32), counts=np.zeros((1, stride), np.float32),
                     state=np.zeros(1, np.int32), stride=stride)
        rng = fba.Rng.replay(np.concatenate([np.array([7, 11], np.uint32), words[pos:pos + nw]]))
        slot = np.zeros(1, np.int64)
        rc = ctx.L.fba_belief_breed_into(dst.h, ptr(slot), 1, donor.h, fc.h, int(g[P + "kind"]), C.byref(rng))
        assert rc == 0, ctx.L.fba_last_error(ctx.h)
        assert rng.exhausted, k
        d = dst.download()
        t, o = sim.structure(int(d["struct_id"][0]))
        np.testing.assert_array_equal(t.reshape(-1), g[P + "t_out"][k])
        np.testing.assert_array_equal(o.reshape(-1), g[P + "o_out"][k])
        assert d["state"][0] == 1                   # breed: the structure donor's domain state
        pos += nw
        donor.free(), dst.free()
    fc.free()
    sim.close()
    ctx.close()


def test_argument_checks_of_the_structure_belief_calls(env):
    """bad arguments come back as FbaError (FBA_ERR_INVALID / _CAPACITY) with the reference's wording where it has one,
    never as a crash or a silent no-op"""
    fba, g, sim = env
    from fba_pomdp_b200 import capi
    from fba_pomdp_b200.capi import ptr
    from fba_pomdp_b200.structure_beliefs import (CheatingReinvigoration, NestedBelief, StructureIncubatorSampling,
                                                  replace_from)
    import ctypes as C
    N, stride = int(g["meta/N"]), int(g["meta/stride"])
    # constructor checks with the reference's messages
    for bad in (lambda: CheatingReinvigoration(0, 1, -1.0), lambda: CheatingReinvigoration(8, 0, -1.0),
                lambda: CheatingReinvigoration(8, 1, 0.0), lambda: StructureIncubatorSampling(0, 1, 0.5, 0),
                lambda: StructureIncubatorSampling(8, 1, 0.0, 0), lambda: StructureIncubatorSampling(8, 1, 1.5, 0),
                lambda: NestedBelief(0, 4), lambda: NestedBelief(4, 0)):
        with pytest.raises(fba.FbaError) as e:
            bad()
        assert e.value.status == capi.ERR_INVALID
    flat = fba.BARejectionSampling(N)
    flat.initiate(sim, **particles(g, "cheat/init_c"), stride=stride)
    wtd = fba.BAImportanceSampling(N)
    wtd.initiate(sim, **particles(g, "cheat/init_b"), stride=stride)
    L, ctx = flat.L, flat.ctx
    rng = fba.Rng.philox(1)
    replace_from(wtd, [], flat, [])                                        # empty: a no-op, not an error
    idx = np.zeros(4, np.int64)
    assert L.fba_belief_least_likely(wtd.h, N, ptr(idx)) == capi.ERR_INVALID          # n must be < N (WeightedFilter.cpp:208)
    assert L.fba_belief_least_likely(flat.h, 2, ptr(idx)) == capi.ERR_INVALID         # flat filters have no weights
    assert L.fba_belief_cheat(flat.h, wtd.h, 2, C.byref(rng)) == capi.ERR_INVALID     # roles swapped
    assert L.fba_belief_promote(flat.h, wtd.h, 0.5, C.byref(rng), None) == capi.ERR_INVALID
    slot = np.array([N], np.int64)
    assert L.fba_belief_breed_into(wtd.h, ptr(slot), 1, flat.h, flat.h, 0, C.byref(rng)) == capi.ERR_INVALID   # slot out of range
    assert L.fba_belief_breed_into(wtd.h, ptr(idx), 0, flat.h, flat.h, 0, C.byref(rng)) == capi.ERR_INVALID    # amount < 1
    assert L.fba_belief_breed_into(flat.h, ptr(idx), 1, wtd.h, flat.h, 0, C.byref(rng)) == capi.ERR_INVALID    # weighted donor
    short = fba.Rng.replay(np.zeros(1, np.uint32))
    assert L.fba_belief_cheat(wtd.h, flat.h, 2, C.byref(short)) == capi.ERR_RNG_UNDERRUN and short.cursor == 0
    # a threshold every shadow particle passes: the reference divides by zero there
    with pytest.raises(fba.FbaError):
        n = C.c_int64(0)
        from fba_pomdp_b200.beliefs import _check
        _check(ctx.h, L.fba_belief_promote(wtd.h, flat.h, 0.5 / N, C.byref(rng), C.byref(n)))
    # history calls
    ln, ac, ob = np.array([2], np.int32), np.array([0, 1], np.int32), np.array([0, 0], np.int32)
    with pytest.raises(fba.FbaError):
        flat.add_history_counts(ln, ac, ob, np.array([0, 1, 10 ** 6], np.int32))       # state out of range
    with pytest.raises(fba.FbaError):
        flat.add_history_counts(ln, np.array([0, 99], np.int32), ob, np.zeros(3, np.int32))   # action out of range
    with pytest.raises(fba.FbaError):
        flat.sample_state_history("rs", np.array([0], np.int32), ac[:0], ob[:0], rng)  # an episode without steps
    nb = NestedBelief(2, 3)
    with pytest.raises(fba.FbaError):
        nb.initiate(sim, struct_id=np.zeros(2, np.int32), counts=g["cheat/init_c_counts"][:2],
                    states=np.full((2, 3), 10 ** 6, np.int32))                          # bottom state out of range
    nb.free()
    flat.free()
    wtd.free()
