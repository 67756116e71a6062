// C++ host driver for MANY RUNS IN LOCKSTEP on one GPU: experiment::bapomdp::run
// (src/experiments/BAPOMDPExperiment.cpp:32-78) with its `num_runs` runs advanced together instead of
// one after the other. Per run everything is what the reference does — prior particles from the
// reference's own prior, one environment of the reference's own kind, episode::run's loop
// (src/experiments/Episode.cpp:16-70): selectAction, env.step, updateEstimation unless terminal,
// discounted return — but the R beliefs live in one fba_runs object and the R planners are R device
// trees: one fba_runs_plan and one fba_runs_update_estimation call per time step serve all runs.
// With sims_per_wave = 1 each run's POMCP search is the sequential algorithm (bit-identical to a
// stand-alone search); the GPU is filled by the number of runs.
//
// Runs do not wait for each other: a run whose episode ends starts its next one at the next global
// step (its belief alone gets resetDomainStateDistribution), so the batch stays full until the runs
// run out of episodes. Only the environment steps (R cheap host calls per global step) and the episode
// bookkeeping stay on the host. Uses the reference's public API + the accessors of INTEGRATION.md §2.
#ifndef FBA_B200_CUDA_EXPERIMENT_HPP
#define FBA_B200_CUDA_EXPERIMENT_HPP

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "CudaBeliefs.hpp"

#include "configurations/BAConf.hpp"
#include "environment/Environment.hpp"
#include "experiments/BAPOMDPExperiment.hpp"
#include "environment/Reward.hpp"
#include "environment/Terminal.hpp"

namespace fba_b200 {

// returns[episode][run] = discounted return of that episode of that run.
// result (may be NULL): the reference's own experiment::bapomdp::Result (BAPOMDPExperiment.hpp:21-37),
// filled the way experiment::bapomdp::run fills it (BAPOMDPExperiment.cpp:61-65) — per episode the
// returns of all runs, in run order, into `ret`, and each run's seconds per step of that episode into
// `duration` — so that Result::log prints the reference's result file (BAPOMDPExperiment.cpp:20-30:
// "return mean, return var, return count, return stder, step duration mean") and the analysis/
// scripts read it unchanged. In lockstep one global step serves every active run at once: its wall
// time is split evenly over them (the reference's boost::timer is CPU time of one run on one core).
inline std::vector<std::vector<double>> runBatchedExperiment(
    BAPOMDP const& bapomdp,
    configurations::BAConf const& conf,
    int runs,
    int sims_per_wave                  = 1,
    uint64_t seed                      = 4711,
    int device                         = 0,
    experiment::bapomdp::Result* result = nullptr)
{
    if (runs < 1) throw std::string("runBatchedExperiment: runs must be at least 1");
    // FBA_B200_TRACE=1: wall time per phase on stderr
    bool const trace = std::getenv("FBA_B200_TRACE") != nullptr;
    double phase_s[5] = {0, 0, 0, 0, 0}; // init, reset, plan, environment, update
    long steps        = 0;
    auto clock0       = std::chrono::steady_clock::now();
    auto lap          = [&](int k) {
        auto const now = std::chrono::steady_clock::now();
        phase_s[k] += std::chrono::duration<double>(now - clock0).count();
        clock0 = now;
    };
    size_t const n = conf.belief_conf.particle_amount;
    if (n < 1) throw "cannot initiate belief with n " + std::to_string(n); // BAImportanceSampling.cpp:19-22
    int const h         = conf.horizon;
    int const sims      = conf.planner_conf.mcts_simulation_amount;
    int const max_depth = conf.planner_conf.mcts_max_depth == -1 ? h : conf.planner_conf.mcts_max_depth;
    double const u = conf.planner_conf.mcts_exploration_const, gamma = conf.discount;

    // domain start states are drawn on the device from the empirical start distribution of the
    // reference's own domain (2^16 host draws, once)
    CudaSimulator cuda(bapomdp, device, 4096, -1, 1 << 16);
    fba_ctx* ctx = cuda.ctx();
    fba_rng rng;
    rng.mode    = FBA_RNG_PHILOX;
    rng.words   = nullptr;
    rng.n_words = rng.cursor = 0;
    rng.seed    = seed;
    rng.offset  = 0;

    // Belief::initiate of every run = runs x n samples of the reference's own prior (host); distinct
    // (structure, count block) pairs are uploaded once as prototypes (as CudaParticleBelief::initiate).
    // A prior that keeps returning the same particle (the tabular priors, the factored ones without a
    // structure prior) is recognised after the first 4096 samples: then every particle IS that
    // prototype and only its domain start state is drawn, on the device.
    size_t const total = (size_t)runs * n;
    size_t const probe = std::min<size_t>(total, 4096);
    std::vector<int32_t> proto(total), proto_sid;
    std::vector<std::vector<float>> proto_blocks;
    std::map<std::string, int32_t> known;
    std::vector<float> block;
    size_t stride = 0, sampled = 0;
    auto sample_prior = [&](size_t upto) {
        for (; sampled < upto; ++sampled)
        {
            auto p            = static_cast<BAState const*>(bapomdp.sampleStartState());
            int32_t const sid = cuda.describe(p, &block);
            std::string key((char const*)&sid, sizeof(sid));
            key.append((char const*)block.data(), block.size() * sizeof(float));
            auto it = known.find(key);
            if (it == known.end())
            {
                it = known.emplace(std::move(key), (int32_t)proto_sid.size()).first;
                proto_sid.push_back(sid);
                proto_blocks.push_back(block);
                stride = std::max(stride, block.size());
            }
            proto[sampled] = it->second;
            bapomdp.releaseState(p);
        }
    };
    sample_prior(probe);
    bool const deterministic_prior = proto_sid.size() == 1;
    if (!deterministic_prior) sample_prior(total);
    struct RunsGuard
    {
        fba_runs* r = nullptr;
        ~RunsGuard() { fba_runs_destroy(r); }
    } batch;
    check(ctx, fba_runs_create(ctx, cuda.model(), runs, (int64_t)n, (int64_t)stride, &batch.r), "fba_runs_create");
    stride = (size_t)fba_belief_stride(fba_runs_belief(batch.r));
    {
        std::vector<float> flat(proto_sid.size() * stride, 0.0f);
        for (size_t k = 0; k < proto_sid.size(); ++k)
            std::copy(proto_blocks[k].begin(), proto_blocks[k].end(), flat.begin() + k * stride);
        if (deterministic_prior)
            check(ctx, fba_runs_init_sampled(batch.r, 1, proto_sid.data(), flat.data(), nullptr, &rng),
                  "fba_runs_init_sampled");
        else
        { // domain states are redrawn at the first episode start anyway (resetDomainStateDistribution)
            std::vector<int32_t> state(total, 0);
            check(ctx,
                  fba_runs_init(batch.r, (int32_t)proto_sid.size(), proto_sid.data(), flat.data(), proto.data(),
                                state.data()),
                  "fba_runs_init");
        }
    }

    // one environment object serves every run (Environment::step is const and keeps the state outside)
    auto const env = factory::makeEnvironment(conf.domain_conf);
    // the domain's own action objects, by index (RBAPOUCT.cpp:74)
    std::vector<Action const*> legal;
    {
        auto p = bapomdp.sampleStartState();
        bapomdp.addLegalActions(p, &legal);
        bapomdp.releaseState(p);
    }
    std::vector<Action const*> action_of((size_t)cuda.A(), nullptr);
    for (auto a : legal)
        if (a->index() >= 0 && a->index() < cuda.A()) action_of[(size_t)a->index()] = a;
    for (auto a : action_of)
        if (!a)
        {
            for (auto l : legal) bapomdp.releaseAction(l);
            throw std::string("runBatchedExperiment: state-dependent action sets are not supported");
        }

    // Every run walks through its own episodes at its own pace (no run waits for the slowest episode
    // of the batch): at each global step a run either starts its next episode — its belief gets
    // resetDomainStateDistribution — or continues the current one, until it has done them all.
    lap(0);
    int const E = conf.num_episodes;
    std::vector<std::vector<double>> returns((size_t)E, std::vector<double>((size_t)runs, 0.0));
    std::vector<State const*> s((size_t)runs, nullptr);
    std::vector<int> episode((size_t)runs, 0), t((size_t)runs, 0);
    std::vector<uint8_t> active((size_t)runs, 1), starting((size_t)runs), updating((size_t)runs);
    std::vector<int32_t> depth((size_t)runs, 0), act((size_t)runs, 0), obs((size_t)runs, 0);
    std::vector<double> disc((size_t)runs, 1.0);
    std::vector<std::vector<double>> seconds((size_t)E, std::vector<double>((size_t)runs, 0.0));
    std::vector<std::vector<int>> length((size_t)E, std::vector<int>((size_t)runs, 0));
    std::vector<int> serving((size_t)runs, -1); // the episode each run's current global step belongs to
    try
    {
        for (;;)
        {
            auto const step_t0 = std::chrono::steady_clock::now();
            bool any = false, any_start = false;
            for (int r = 0; r < runs; ++r)
            {
                starting[(size_t)r] = 0;
                if (!active[(size_t)r]) continue;
                any = true;
                if (!s[(size_t)r])
                { // episode start (BAPOMDPExperiment.cpp:54-59, Episode.cpp:31)
                    s[(size_t)r]        = env->sampleStartState();
                    t[(size_t)r]        = 0;
                    disc[(size_t)r]     = 1.0; // Discount(_discount) starts at 1 (Discount.cpp:3-6)
                    starting[(size_t)r] = 1;
                    any_start           = true;
                }
                depth[(size_t)r] = std::min(h - t[(size_t)r], max_depth); // RBAPOUCT.cpp:80
            }
            if (!any) break;
            ++steps;
            lap(3);
            if (any_start)
                check(ctx, fba_runs_reset_domain_states(batch.r, starting.data(), &rng), "fba_runs_reset_domain_states");
            lap(1);
            check(ctx,
                  fba_runs_plan(batch.r, sims, depth.data(), u, gamma, sims_per_wave, active.data(), &rng, act.data(),
                                nullptr, nullptr),
                  "fba_runs_plan");
            lap(2);
            for (int r = 0; r < runs; ++r)
            {
                updating[(size_t)r] = 0;
                serving[(size_t)r]  = -1;
                if (!active[(size_t)r]) continue;
                serving[(size_t)r] = episode[(size_t)r];
                ++length[(size_t)episode[(size_t)r]][(size_t)r];
                Observation const* o(nullptr);
                Reward rew(0);
                auto const terminal = env->step(&s[(size_t)r], action_of[(size_t)act[(size_t)r]], &o, &rew);
                obs[(size_t)r]      = o->index();
                env->releaseObservation(o);
                returns[(size_t)episode[(size_t)r]][(size_t)r] += rew.toDouble() * disc[(size_t)r]; // Return::add
                disc[(size_t)r] *= gamma;                                                           // Discount::increment
                ++t[(size_t)r];
                if (!terminal.terminated()) updating[(size_t)r] = 1; // no belief update after a terminal step
                if (terminal.terminated() || t[(size_t)r] >= h)
                { // episode over: the next global step starts this run's next episode, if any
                    env->releaseState(s[(size_t)r]);
                    s[(size_t)r] = nullptr;
                    if (++episode[(size_t)r] >= E) active[(size_t)r] = 0;
                }
            }
            lap(3);
            check(ctx, fba_runs_update_estimation(batch.r, act.data(), obs.data(), updating.data(), &rng, nullptr),
                  "fba_runs_update_estimation");
            lap(4);
            if (result)
            { // the update call above synchronises only when it returns likelihoods: wait for the step's work
                check(ctx, fba_ctx_synchronize(ctx), "fba_ctx_synchronize");
                int served = 0;
                for (int r = 0; r < runs; ++r) served += serving[(size_t)r] >= 0;
                double const dt =
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - step_t0).count() / std::max(served, 1);
                for (int r = 0; r < runs; ++r)
                    if (serving[(size_t)r] >= 0) seconds[(size_t)serving[(size_t)r]][(size_t)r] += dt;
            }
        }
        if (trace)
            std::fprintf(stderr,
                         "runBatchedExperiment: %ld global steps; init %.3f s, reset %.3f s, plan %.3f s, environment "
                         "%.3f s, update %.3f s\n",
                         steps, phase_s[0], phase_s[1], phase_s[2], phase_s[3], phase_s[4]);
    } catch (...)
    {
        for (auto p : s)
            if (p) env->releaseState(p);
        for (auto l : legal) bapomdp.releaseAction(l);
        throw;
    }
    for (auto l : legal) bapomdp.releaseAction(l);
    if (result)
    {
        *result = experiment::bapomdp::Result(E);
        for (int e = 0; e < E; ++e)
            for (int r = 0; r < runs; ++r)
            { // BAPOMDPExperiment.cpp:63-64
                result->r[(size_t)e].ret.add(returns[(size_t)e][(size_t)r]);
                result->r[(size_t)e].duration.add(seconds[(size_t)e][(size_t)r] / std::max(length[(size_t)e][(size_t)r], 1));
            }
    }
    return returns;
}

} // namespace fba_b200

#endif // FBA_B200_CUDA_EXPERIMENT_HPP
