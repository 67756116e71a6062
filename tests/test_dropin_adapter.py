"""GPU: the C++ host adapters (fba-pomdp_b200/host/CudaBeliefs.hpp) dropped into the reference's OWN
episode loop and POMCP planner (oracle/ref_harness.cpp:ref_adapter_episodes, compiled against the
unmodified reference). The adapters run in PHILOX mode, so the check is statistical: episode returns
under the CUDA belief agree with returns under the reference's CPU belief within 4 standard errors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

pyref = pytest.importorskip("pyref")
if not pyref.available():
    pytest.skip("oracle/_ref/libfba_ref.so not built", allow_module_level=True)

CASES = [
    # domain, kwargs, belief kinds (reference, cuda), particles, episodes
    ("episodic-tiger", dict(), (0, 1), 256, 80),
    ("episodic-tiger", dict(), (2, 3), 256, 80),
    ("episodic-factored-tiger", dict(size=3, factored=True), (0, 1), 128, 40),
    ("centered-collision-avoidance", dict(size=1, width=3, height=3, factored=True), (0, 1), 128, 40),
    ("linear-sysadmin", dict(size=3, factored=True), (0, 1), 64, 20),
    ("gridworld", dict(size=3), (0, 1), 16, 10),
    # gridworld size 5: the adapter switches to base+delta storage by itself (dense block = 720 KB)
    ("gridworld", dict(size=5), (0, 1), 16, 6),
    ("gridworld", dict(size=5), (2, 3), 16, 6),
    # reinvigoration: the reference's ReinvigoratingRejectionSampling vs the CUDA adapter
    ("episodic-factored-tiger", dict(size=3, factored=True, structure_prior="match-uniform"), (4, 5), 128, 40),
    ("centered-collision-avoidance", dict(size=1, width=3, height=3, factored=True,
                                          structure_prior="match-uniform"), (4, 5), 128, 40),
    ("linear-sysadmin", dict(size=3, factored=True), (4, 5), 64, 20),
    # MHNIPS2018 (threshold -4: MH runs every few steps): the reference's class vs fba_b200::CudaMHNIPS2018
    ("episodic-factored-tiger", dict(size=3, factored=True, structure_prior="match-uniform"), (6, 7), 96, 40),
    ("centered-collision-avoidance", dict(size=1, width=3, height=3, factored=True,
                                          structure_prior="match-uniform"), (6, 7), 64, 30),
    # the composite structure beliefs, settings of the reference's own integration tests (test/test.cpp:296-353):
    # CheatingReinvigoration (threshold -3) and StructureIncubatorSampling (threshold .05) vs the CUDA adapters
    ("episodic-factored-tiger", dict(size=3, factored=True, structure_prior="match-uniform"), (8, 9), 96, 40),
    ("centered-collision-avoidance", dict(size=1, width=3, height=3, factored=True,
                                          structure_prior="match-uniform"), (8, 9), 64, 30),
    ("episodic-factored-tiger", dict(size=3, factored=True, structure_prior="match-uniform"), (10, 11), 96, 40),
    ("centered-collision-avoidance", dict(size=1, width=3, height=3, factored=True,
                                          structure_prior="match-uniform"), (10, 11), 64, 30),
    ("linear-sysadmin", dict(size=3, factored=True), (10, 11), 64, 20),
    # NestedBelief, n top particles with n^2 bottom states each as the factory sizes it (BABelief.cpp:66-69)
    ("episodic-tiger", dict(), (12, 13), 10, 60),
    ("episodic-factored-tiger", dict(size=3, factored=True, structure_prior="match-uniform"), (12, 13), 8, 40),
    ("gridworld", dict(size=3), (12, 13), 6, 10),
    ("episodic-tiger", dict(sampled=True), (12, 13), 10, 60),
    # MHwithinGibbs (threshold -4: the Gibbs chain runs every few steps), message passing and rejection sampling
    ("episodic-factored-tiger", dict(size=3, factored=True, structure_prior="match-uniform"), (14, 15), 64, 40),
    ("episodic-factored-tiger", dict(size=3, factored=True, structure_prior="match-uniform"), (16, 17), 64, 40),
    ("centered-collision-avoidance", dict(size=1, width=3, height=3, factored=True,
                                          structure_prior="match-uniform"), (14, 15), 48, 30),
    # --dirichlet_sampling_method regular: the adapter reads the mode from the simulator
    ("episodic-tiger", dict(sampled=True), (0, 1), 256, 80),
    ("episodic-factored-tiger", dict(size=3, factored=True, sampled=True), (0, 1), 128, 40),
]


@pytest.mark.parametrize("domain,kw,kinds,n,episodes", CASES)
def test_reference_episode_loop_with_cuda_belief(domain, kw, kinds, n, episodes):
    horizon = 8
    r = pyref.Ref(domain, horizon=horizon, seed="7", **kw)
    try:
        ref = r.adapter_episodes(kinds[0], n, "po-uct", 48, episodes)
        ours = r.adapter_episodes(kinds[1], n, "po-uct", 48, episodes)
        events = r.adapter_events()
    finally:
        r.close()
    if kinds[1] in (7, 9, 15, 17):      # MH runs / cheats / Gibbs chains did happen during these episodes
        assert events >= 2, events
    assert np.all(np.isfinite(ours))
    se = np.sqrt(ref.var(ddof=1) / len(ref) + ours.var(ddof=1) / len(ours)) + 1e-9
    assert abs(ref.mean() - ours.mean()) <= 4.0 * se + 1e-6, (ref.mean(), ours.mean(), se)


PLANNER_CASES = [
    # domain, kwargs, particles, episodes, simulations, wave
    ("episodic-tiger", dict(), 512, 150, 128, 8),
    ("gridworld", dict(size=3), 64, 40, 128, 16),
    ("gridworld", dict(size=5), 16, 8, 64, 16),
    ("centered-collision-avoidance", dict(size=1, width=3, height=3, factored=True), 128, 60, 128, 16),
    ("linear-sysadmin", dict(size=3, factored=True), 64, 30, 128, 16),
]


@pytest.mark.parametrize("planner", ["cuda-po-uct", "cuda-tree-po-uct"])
@pytest.mark.parametrize("domain,kw,n,episodes,sims,wave", PLANNER_CASES)
def test_reference_episode_loop_with_cuda_planner(domain, kw, n, episodes, sims, wave, planner):
    """fba_b200::CudaBatchedPOUCT (wave-parallel POMCP, tree on the host, simulator steps and leaf
    rollouts batched on the GPU) and fba_b200::CudaTreePOUCT (the tree itself on the device, whole
    simulations inside one kernel) + the CUDA belief, inside the reference's own episode::run, against the reference's
    RBAPOUCT + BAImportanceSampling: mean episode returns agree within 4 standard errors, and the
    planner is clearly better than acting at random."""
    horizon = 8
    r = pyref.Ref(domain, horizon=horizon, seed="11", **kw)
    try:
        ref = r.adapter_episodes(0, n, "po-uct", sims, episodes)
        ours = r.adapter_episodes(1, n, "%s:%d" % (planner, wave), sims, episodes)
        rand = r.adapter_episodes(0, n, "random", sims, episodes)
    finally:
        r.close()
    assert np.all(np.isfinite(ours))
    se = np.sqrt(ref.var(ddof=1) / len(ref) + ours.var(ddof=1) / len(ours)) + 1e-9
    assert abs(ref.mean() - ours.mean()) <= 4.0 * se + 1e-6, (ref.mean(), ours.mean(), se)
    if domain in ("episodic-tiger", "centered-collision-avoidance"):
        # planning matters in these two at this horizon: the CUDA planner is not worse than acting at
        # random (collision returns are -950 a piece, so allow two standard errors)
        se_r = np.sqrt(rand.var(ddof=1) / len(rand) + ours.var(ddof=1) / len(ours))
        assert ours.mean() > rand.mean() - 2.0 * se_r, (ours.mean(), ref.mean(), rand.mean())


BATCH_CASES = [
    # domain, kwargs, particles, runs, episodes, simulations
    ("episodic-tiger", dict(), 256, 48, 3, 128),
    ("episodic-factored-tiger", dict(size=3, factored=True), 128, 32, 2, 96),
    ("linear-sysadmin", dict(size=3, factored=True), 64, 32, 2, 96),
    ("gridworld", dict(size=3), 32, 24, 2, 96),
    # a prior that samples a structure per particle: the driver takes the exact per-particle path
    ("episodic-factored-tiger", dict(size=3, factored=True, structure_prior="match-uniform"), 64, 24, 2, 64),
]


@pytest.mark.parametrize("domain,kw,n,runs,episodes,sims", BATCH_CASES)
def test_many_runs_in_lockstep_match_the_reference_experiment(domain, kw, n, runs, episodes, sims):
    """fba_b200::runBatchedExperiment (host/CudaExperiment.hpp): experiment::bapomdp::run with its runs
    advanced TOGETHER — all beliefs in one fba_runs object, all planners device trees, one plan and
    one update call per time step for every run — against the reference's own experiment loop run
    once per run (BAImportanceSampling + RBAPOUCT on the CPU): per-episode mean returns over the runs
    agree within 4 standard errors."""
    horizon = 8
    r = pyref.Ref(domain, horizon=horizon, seed="31", **kw)
    try:
        ours, _ = r.batched_episodes(n, runs, sims, episodes)
        ref = np.stack([r.adapter_episodes(0, n, "po-uct", sims, episodes) for _ in range(runs)], axis=1)
    finally:
        r.close()
    assert ours.shape == ref.shape == (episodes, runs) and np.all(np.isfinite(ours))
    for e in range(episodes):
        se = np.sqrt(ref[e].var(ddof=1) / runs + ours[e].var(ddof=1) / runs) + 1e-9
        assert abs(ref[e].mean() - ours[e].mean()) <= 4.0 * se + 1e-6, (e, ref[e].mean(), ours[e].mean(), se)


INIT_CASES = [
    # domain, kwargs, particles, reference draws, bound on the host prior samples (None: all of them)
    ("linear-sysadmin", dict(size=10, factored=True), 1_000_000, 20_000, 4100),          # config 5: one prototype
    ("gridworld", dict(size=5), 200_000, 300, 4),                                        # 720 KB tables, one prototype
    ("centered-collision-avoidance", dict(size=1, width=5, height=5, factored=True,
                                          structure_prior="match-uniform"), 400_000, 100_000, 70_000),
    ("episodic-factored-tiger", dict(size=3, factored=True, structure_prior="match-uniform"), 3000, 3000, None),
]


@pytest.mark.parametrize("domain,kw,n,n_ref,host_bound", INIT_CASES)
def test_initiate_matches_the_reference_prior_and_is_fast(domain, kw, n, n_ref, host_bound):
    """CudaParticleBelief::initiate (host/CudaBeliefs.hpp) against the reference's own
    BAImportanceSampling::initiate = N x sampleStartState (BAImportanceSampling.cpp:49-60): the
    distributions of the domain start state and of the prior's structure agree (per category, 5 standard
    errors of the two-sample difference, plus the estimation error of the prototype frequencies where the
    adapter stopped sampling the prior early), and large beliefs take seconds: 10^6 sysadmin-10 particles
    in 3-5 s (asserted: under 20 s) where 10^6 reference prior samples take about a minute."""
    r = pyref.Ref(domain, horizon=8, seed="5", **kw)
    try:
        res = r.adapter_initiate(n, n_ref)
    finally:
        r.close()
    m = res["host_samples"]
    assert m == n if host_bound is None else m <= host_bound, m
    if n >= 1_000_000:
        # 2.8 s typical; the bound is about "seconds, not the minute 10^6 reference prior samples
        # take" and leaves room for a slow or busy host — a timing assertion must not make the suite flaky
        assert res["seconds"] < 20.0, res["seconds"]
    for key in ("state", "sid"):
        c, f = res["cuda_" + key], res["ref_" + key]
        k = int(max(c.max(), f.max())) + 1
        pc, pf = np.bincount(c, minlength=k) / len(c), np.bincount(f, minlength=k) / len(f)
        p = (pc * len(c) + pf * len(f)) / (len(c) + len(f))
        var = p * (1 - p) * (1.0 / len(c) + 1.0 / len(f))
        if key == "sid" and m < n:
            var = var + p * (1 - p) / m            # prototype frequencies were estimated from m prior samples
        assert np.all(np.abs(pc - pf) <= 5.0 * np.sqrt(var) + 1e-12), (key, np.abs(pc - pf).max())
    print("\ninitiate %s: %d particles in %.2f s, %d host prior samples" % (domain, n, res["seconds"], m))


def test_batched_experiment_writes_the_reference_result_file(tmp_path):
    """runBatchedExperiment fills the reference's own experiment::bapomdp::Result and Result::log writes
    the reference's result file (BAPOMDPExperiment.cpp:20-30,61-65): read back the way
    analysis/preprocess/merge_result_files.py:38 does (np.loadtxt with ',' after two '#' header lines),
    one row per episode = return mean, sample variance, count = runs, standard error, mean seconds per
    step — consistent with the raw returns; two such files merge with that script's formula."""
    runs, episodes = 32, 3
    r = pyref.Ref("episodic-tiger", horizon=8, seed="3")
    try:
        files, rets = [], []
        for k, seed in enumerate((4711, 9001)):
            path = tmp_path / ("%d.res" % k)
            ret, dt = r.batched_experiment_file(128, runs, 64, episodes, path, seed=seed)
            files.append(path)
            rets.append(ret)
    finally:
        r.close()
    head = open(files[0]).read().splitlines()[:2]
    assert head == ["# version 1:",
                    "# return mean, return var, return count, return stder, step duration mean"]
    tabs = [np.loadtxt(f, delimiter=",") for f in files]
    for tab, ret in zip(tabs, rets):
        assert tab.shape == (episodes, 5)
        np.testing.assert_allclose(tab[:, 0], ret.mean(1), rtol=1e-5)          # 6 significant digits in the file
        np.testing.assert_allclose(tab[:, 1], ret.var(1, ddof=1), rtol=1e-4)
        np.testing.assert_array_equal(tab[:, 2], runs)
        np.testing.assert_allclose(tab[:, 3], np.sqrt(ret.var(1, ddof=1) / runs), rtol=1e-4)
        assert np.all(tab[:, 4] > 0) and np.all(tab[:, 4] < 1.0)
    # merge_result_files.py:63-80 on the two files = the statistics of the pooled runs
    n = tabs[0][:, 2] + tabs[1][:, 2]
    mu = (tabs[0][:, 0] * tabs[0][:, 2] + tabs[1][:, 0] * tabs[1][:, 2]) / n
    np.testing.assert_allclose(mu, np.concatenate(rets, 1).mean(1), rtol=1e-5)
