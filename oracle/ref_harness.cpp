// TEST INFRASTRUCTURE — not shipped, never on the product path.
//
// A thin extern "C" window onto the UNMODIFIED reference (samkatt/fba-pomdp), compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libfba_ref.so. It exists to
//   (1) pin oracle/fba_oracle.c (the plain-C restatement) against the reference's real classes,
//   (2) generate the golden fixtures under tests/golden/ (oracle/gen_golden.py), and
//   (3) be the CPU baseline (`bench.py --impl reference`, cpu_baseline.kind = "reference").
//
// It drives the reference exactly as its own callers do:
//   beliefs::BAImportanceSampling::{initiate,updateEstimation,resetDomainStateDistribution}
//       (src/beliefs/bayes-adaptive/BAImportanceSampling.cpp:49-111)
//   beliefs::importance_sampling::{update,resample}
//       (src/beliefs/particle_filters/ImportanceSampler.hpp:31-94)
//   beliefs::BARejectionSampling / rejectSample (src/beliefs/particle_filters/RejectionSampling.hpp:26-72)
//   beliefs::bayes_adaptive::factored::ReinvigoratingRejectionSampling
//       (src/beliefs/bayes-adaptive/factored/ReinvigoratingRejectionSampling.cpp:89-131)
//   planners::RBAPOUCT::rollout (src/planners/bayes-adaptive/RBAPOUCT.cpp:295-323)
// Config structs are filled programmatically, as test/test.cpp:170-179 does.
//
// RNG tap: the reference draws from one global std::mt19937 (src/utils/random.cpp:11). The harness
// never replaces it; it copies the engine before an operation ("mark") and afterwards advances the
// copy until it equals the live engine, which yields the exact 32-bit words the operation consumed.
// Those words are the replay stream fed to oracle/fba_oracle.c and to the CUDA path.
//
// Built with -fno-access-control (this TU only) so private filters can be dumped.

#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "easylogging++.h"

#include "bayes-adaptive/models/factored/FBAPOMDP.hpp"
#include "bayes-adaptive/models/table/BADomainExtension.hpp"
#include "bayes-adaptive/models/table/BAPOMDP.hpp"
#include "bayes-adaptive/states/factored/BABNModel.hpp"
#include "bayes-adaptive/states/factored/FBAPOMDPState.hpp"
#include "bayes-adaptive/states/table/BAPOMDPState.hpp"
#include "beliefs/bayes-adaptive/BAImportanceSampling.hpp"
#include "beliefs/bayes-adaptive/BARejectionSampling.hpp"
#include "beliefs/bayes-adaptive/NestedBelief.hpp"
#include "beliefs/bayes-adaptive/factored/MHNIPS2018.hpp"
#include "beliefs/bayes-adaptive/factored/MHwithinGibbs.hpp"
#include "beliefs/bayes-adaptive/factored/StructureIncubatorSampling.hpp"
#include "beliefs/bayes-adaptive/prototypes/CheatingReinvigoration.hpp"
#include "beliefs/bayes-adaptive/factored/ReinvigoratingRejectionSampling.hpp"
#include "beliefs/particle_filters/ImportanceSampler.hpp"
#include "beliefs/particle_filters/RejectionSampling.hpp"
#include "configurations/FBAConf.hpp"
#include "domains/POMDP.hpp"
#include "domains/collision-avoidance/CollisionAvoidance.hpp"
#include "domains/gridworld/GridWorld.hpp"
#include "domains/sysadmin/SysAdmin.hpp"
#include "environment/Action.hpp"
#include "environment/Environment.hpp"
#include "environment/Observation.hpp"
#include "environment/Reward.hpp"
#include "environment/State.hpp"
#include "environment/Terminal.hpp"
#include "planners/bayes-adaptive/RBAPOUCT.hpp"
#include "utils/random.hpp"

// the C++ host adapters of this repo (the drop-in beliefs), compiled against the reference here
#define FBA_B200_PRIVATE_ACCESS
#include "CudaBeliefs.hpp"
#include "CudaExperiment.hpp"
#include "CudaMH.hpp"
#include "CudaPlanner.hpp"
#include "CudaStructureBeliefs.hpp"
#include "environment/Discount.hpp"
#include "environment/Horizon.hpp"
#include "environment/Return.hpp"
#include "experiments/Episode.hpp"
#include "planners/Planner.hpp"

INITIALIZE_EASYLOGGINGPP

namespace {

using FactoredPOMDP = ::bayes_adaptive::factored::FBAPOMDP;
using ReinvRS  = ::beliefs::bayes_adaptive::factored::ReinvigoratingRejectionSampling;
using RefMH    = ::beliefs::bayes_adaptive::factored::MHNIPS2018;
using RefCheat = ::beliefs::bayes_adaptive::prototypes::CheatingReinvigoration;
using RefIncub = ::beliefs::bayes_adaptive::factored::StructureIncubatorSampling;
using RefGibbs = ::beliefs::bayes_adaptive::factored::MHwithinGibbs;
using RefNested = ::beliefs::bayes_adaptive::NestedBelief;

struct Handle
{
    configurations::FBAConf conf;
    std::unique_ptr<BAPOMDP> sim; // hyper-state simulator (tabular BAPOMDP or FactoredPOMDP)
    std::unique_ptr<Environment> env; // the true domain, for (a,o) scripts
    bool factored = false;
    std::vector<int> feat_s, feat_o;

    std::unique_ptr<beliefs::BAImportanceSampling> is_belief;
    std::unique_ptr<beliefs::BARejectionSampling> rs_belief;
    std::unique_ptr<ReinvRS> reinv_belief;
    std::unique_ptr<RefMH> mh_belief;
    std::unique_ptr<RefCheat> cheat_belief;
    std::unique_ptr<RefIncub> incub_belief;
    std::unique_ptr<RefGibbs> gibbs_belief;
    std::unique_ptr<RefNested> nested_belief;
    std::unique_ptr<planners::RBAPOUCT> planner;

    std::mt19937 mark;
    std::string err;
    long adapter_events = -1; // what the last ref_adapter_episodes' CUDA belief did beyond plain filtering
    double update_seconds = 0; // wall time inside updateEstimation during the last ref_adapter_episodes
    long update_calls     = 0;
    std::vector<double> update_times; // ... per call
};

// forwards to a belief and times its updateEstimation calls
class TimedBelief : public beliefs::BABelief
{
public:
    TimedBelief(beliefs::BABelief* b, Handle* h) : _b(b), _h(h) {}
    void initiate(POMDP const& d) override { _b->initiate(d); }
    void free(POMDP const& d) override { _b->free(d); }
    State const* sample() const override { return _b->sample(); }
    void resetDomainStateDistribution(BAPOMDP const& b) override { _b->resetDomainStateDistribution(b); }
    void updateEstimation(Action const* a, Observation const* o, POMDP const& d) override
    {
        auto t0 = std::chrono::steady_clock::now();
        _b->updateEstimation(a, o, d);
        double const dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        _h->update_seconds += dt;
        _h->update_times.push_back(dt);
        ++_h->update_calls;
    }

private:
    beliefs::BABelief* _b;
    Handle* _h;
};

bool g_rng_initiated = false;

// which particle container a call refers to
enum Filter {
    F_IS = 0, F_RS = 1, F_REINV = 2, F_REINV_FC = 3, F_MH = 4,
    F_CHEAT = 5, F_CHEAT_CORRECT = 6,            // CheatingReinvigoration: weighted belief, flat correct-structure filter
    F_INC = 7, F_INC_FC = 8, F_INC_SHADOW = 9,   // StructureIncubatorSampling: belief, fully connected, weighted shadow
    F_GIBBS = 10, F_NESTED = 11                  // MHwithinGibbs (weighted), NestedBelief's top filter (weighted)
};

BAState const* particleOf(Handle* h, int filter, long i)
{
    switch (filter)
    {
        case F_IS:
            return static_cast<BAState const*>(h->is_belief->_filter.particle(i)->particle);
        case F_RS: return static_cast<BAState const*>(h->rs_belief->_filter.particles()[i]);
        case F_REINV: return h->reinv_belief->_belief.particles()[i];
        case F_REINV_FC: return h->reinv_belief->_fully_connected_belief.particles()[i];
        case F_MH: return h->mh_belief->_belief.particle(i)->particle;
        case F_CHEAT: return h->cheat_belief->_belief.particle(i)->particle;
        case F_CHEAT_CORRECT: return h->cheat_belief->_correct_structured_belief.particles()[i];
        case F_INC: return h->incub_belief->_belief.particles()[i];
        case F_INC_FC: return h->incub_belief->_fully_connected_belief.particles()[i];
        case F_INC_SHADOW: return h->incub_belief->_shadow_belief.particle(i)->particle;
        case F_GIBBS: return h->gibbs_belief->_belief.particle(i)->particle;
        case F_NESTED: return h->nested_belief->_filter.particle(i)->particle.first;
    }
    return nullptr;
}

long filterSize(Handle* h, int filter)
{
    switch (filter)
    {
        case F_IS: return h->is_belief ? (long)h->is_belief->_filter.size() : 0;
        case F_RS: return h->rs_belief ? (long)h->rs_belief->_filter.size() : 0;
        case F_REINV: return h->reinv_belief ? (long)h->reinv_belief->_belief.size() : 0;
        case F_REINV_FC:
            return h->reinv_belief ? (long)h->reinv_belief->_fully_connected_belief.size() : 0;
        case F_MH: return h->mh_belief ? (long)h->mh_belief->_belief.size() : 0;
        case F_CHEAT: return h->cheat_belief ? (long)h->cheat_belief->_belief.size() : 0;
        case F_CHEAT_CORRECT: return h->cheat_belief ? (long)h->cheat_belief->_correct_structured_belief.size() : 0;
        case F_INC: return h->incub_belief ? (long)h->incub_belief->_belief.size() : 0;
        case F_INC_FC: return h->incub_belief ? (long)h->incub_belief->_fully_connected_belief.size() : 0;
        case F_INC_SHADOW: return h->incub_belief ? (long)h->incub_belief->_shadow_belief.size() : 0;
        case F_GIBBS: return h->gibbs_belief ? (long)h->gibbs_belief->_belief.size() : 0;
        case F_NESTED: return h->nested_belief ? (long)h->nested_belief->_filter.size() : 0;
    }
    return 0;
}

uint32_t maskOf(std::vector<int> const& parents)
{
    uint32_t m = 0;
    for (auto p : parents) m |= (1u << p);
    return m;
}

} // namespace

extern "C" {

// domain: reference -D string. factored: 0 = tabular BA-POMDP (makeTBAPOMDP), 1 = FBA-POMDP.
// sampled: 0 = expected Dirichlets (the reference default), 1 = --dirichlet_sampling_method regular
void* ref_open_ex(
    char const* domain,
    int size,
    int width,
    int height,
    int factored,
    char const* structure_prior,
    double discount,
    int horizon,
    char const* seed,
    int sampled)
{
    auto h = new Handle();
    try
    {
        if (!g_rng_initiated)
        {
            rnd::initiate(); // ziggurat tables (unused in expected mode) + time seed, then:
            g_rng_initiated = true;
        }
        std::string seed_str(seed);
        rnd::seed(seed_str);

        auto& c                 = h->conf;
        c.domain_conf.domain    = domain;
        c.domain_conf.size      = size;
        c.domain_conf.width     = width;
        c.domain_conf.height    = height;
        c.structure_prior       = structure_prior;
        c.discount              = discount;
        c.horizon               = horizon;
        c.bayes_sample_method   = sampled ? rnd::sample::Dir::Regular
                                          : rnd::sample::Dir::Expected; // the default (BAConf.hpp:22)
        c.planner_conf.mcts_max_depth         = horizon;
        c.planner_conf.mcts_simulation_amount = 16;

        h->factored = factored != 0;
        h->sim      = factored ? factory::makeFBAPOMDP(c) : factory::makeTBAPOMDP(c);
        h->env      = factory::makeEnvironment(c.domain_conf);

        if (factored)
        {
            auto fs   = static_cast<FactoredPOMDP const*>(h->sim.get())->domainFeatureSize();
            h->feat_s = fs->_S;
            h->feat_o = fs->_O;
        } else
        {
            h->feat_s = {h->sim->domainSize()->_S};
            h->feat_o = {h->sim->domainSize()->_O};
        }
        h->mark = rnd::rng();
    } catch (std::string const& e)
    {
        h->err = e;
    } catch (char const* e)
    {
        h->err = e;
    } catch (std::exception const& e)
    {
        h->err = e.what();
    }
    return h;
}

void* ref_open(
    char const* domain,
    int size,
    int width,
    int height,
    int factored,
    char const* structure_prior,
    double discount,
    int horizon,
    char const* seed)
{
    return ref_open_ex(domain, size, width, height, factored, structure_prior, discount, horizon, seed, 0);
}

char const* ref_error(void* hv)
{
    return static_cast<Handle*>(hv)->err.c_str();
}

void ref_close(void* hv)
{
    auto h = static_cast<Handle*>(hv);
    if (h->sim)
    {
        if (h->is_belief) h->is_belief->free(*h->sim);
        if (h->rs_belief) h->rs_belief->free(*h->sim);
        if (h->reinv_belief) h->reinv_belief->free(*h->sim);
        if (h->mh_belief) h->mh_belief->free(*h->sim);
        if (h->cheat_belief) h->cheat_belief->free(*h->sim);
        if (h->incub_belief) h->incub_belief->free(*h->sim);
        if (h->gibbs_belief) h->gibbs_belief->free(*h->sim);
        if (h->nested_belief) h->nested_belief->free(*h->sim);
    }
    delete h;
}

void ref_reseed(void* hv, char const* seed)
{
    std::string s(seed);
    rnd::seed(s);
    static_cast<Handle*>(hv)->mark = rnd::rng();
}

// out[0..4] = S, A, O, F_S, F_O
void ref_sizes(void* hv, int* out)
{
    auto h = static_cast<Handle*>(hv);
    out[0] = h->sim->domainSize()->_S;
    out[1] = h->sim->domainSize()->_A;
    out[2] = h->sim->domainSize()->_O;
    out[3] = (int)h->feat_s.size();
    out[4] = (int)h->feat_o.size();
}

void ref_feature_sizes(void* hv, int* fs, int* fo)
{
    auto h = static_cast<Handle*>(hv);
    for (size_t i = 0; i < h->feat_s.size(); ++i) fs[i] = h->feat_s[i];
    for (size_t i = 0; i < h->feat_o.size(); ++i) fo[i] = h->feat_o[i];
}

/**** RNG word tap ****/
void ref_rng_mark(void* hv)
{
    static_cast<Handle*>(hv)->mark = rnd::rng();
}

// words the global mt19937 produced since the last mark; returns the count (copies at most cap).
// Gives up (returns -1) after `limit` words without meeting the live engine.
long ref_rng_words_since_mark(void* hv, uint32_t* out, long cap, long limit)
{
    auto h    = static_cast<Handle*>(hv);
    auto twin = h->mark;
    long n    = 0;
    while (!(twin == rnd::rng()))
    {
        auto w = static_cast<uint32_t>(twin());
        if (n < cap && out) out[n] = w;
        ++n;
        if (n > limit) return -1;
    }
    return n;
}

// raw words straight from the live engine (advances it): used to build streams for the oracle
void ref_rng_draw_words(uint32_t* out, long n)
{
    for (long i = 0; i < n; ++i) out[i] = static_cast<uint32_t>(rnd::rng()());
}

/**** beliefs ****/
// kind: F_IS / F_RS / F_REINV. initiate() follows the reference (sampleStartState per particle).
int ref_belief_init(void* hv, int kind, long n, long resample_amount)
{
    auto h = static_cast<Handle*>(hv);
    try
    {
        if (kind == F_IS)
        {
            if (h->is_belief) h->is_belief->free(*h->sim);
            h->is_belief.reset(new beliefs::BAImportanceSampling(n));
            h->is_belief->initiate(*h->sim);
        } else if (kind == F_RS)
        {
            if (h->rs_belief) h->rs_belief->free(*h->sim);
            h->rs_belief.reset(new beliefs::BARejectionSampling(n));
            h->rs_belief->initiate(*h->sim);
        } else
        {
            if (h->reinv_belief) h->reinv_belief->free(*h->sim);
            h->reinv_belief.reset(new ReinvRS(n, resample_amount));
            h->reinv_belief->initiate(*h->sim);
        }
    } catch (std::string const& e)
    {
        h->err = e;
        return 1;
    } catch (char const* e)
    {
        h->err = e;
        return 1;
    }
    return 0;
}

long ref_filter_size(void* hv, int filter)
{
    return filterSize(static_cast<Handle*>(hv), filter);
}

void ref_filter_states(void* hv, int filter, int* out)
{
    auto h = static_cast<Handle*>(hv);
    auto n = filterSize(h, filter);
    for (long i = 0; i < n; ++i)
        out[i] = (filter == F_NESTED) ? -1 : particleOf(h, filter, i)->_domain_state->index();
}

// weights of the importance-sampling filter + its _total_weight
void ref_is_weights(void* hv, double* w, double* total)
{
    auto h = static_cast<Handle*>(hv);
    auto n = filterSize(h, F_IS);
    for (long i = 0; i < n; ++i) w[i] = h->is_belief->_filter.particle(i)->w;
    *total = h->is_belief->_filter._total_weight;
}

// number of float count cells particle i owns (its structure's total CPT size)
long ref_particle_num_counts(void* hv, int filter, long i)
{
    auto h = static_cast<Handle*>(hv);
    auto p = particleOf(h, filter, i);
    auto S = h->sim->domainSize()->_S, A = h->sim->domainSize()->_A, O = h->sim->domainSize()->_O;
    if (!h->factored) return (long)A * S * S + (long)A * S * O;

    auto m = static_cast<FBAPOMDPState const*>(p)->model();
    long n = 0;
    for (auto const& node : m->copyT()) n += (long)node.numParams();
    for (auto const& node : m->copyO()) n += (long)node.numParams();
    return n;
}

// Particle dump in this repo's layout: for a in [0,A): T nodes f = 0..F_S-1, then O nodes
// g = 0..F_O-1; each node's CPT row-major [parent configuration][output].
// Tabular: T(a) = phi[s][a][s'] as [s][s'], O(a) = psi[a][s'][o] as [s'][o]; parent masks = 1.
// t_par: [A*F_S] parent bitmasks, o_par: [A*F_O].
void ref_particle_dump(void* hv, int filter, long i, uint32_t* t_par, uint32_t* o_par, float* counts)
{
    auto h = static_cast<Handle*>(hv);
    auto p = particleOf(h, filter, i);
    auto S = h->sim->domainSize()->_S, A = h->sim->domainSize()->_A, O = h->sim->domainSize()->_O;
    long k = 0;

    if (!h->factored)
    {
        auto const& m = *static_cast<BAPOMDPState const*>(p)->model();
        for (int a = 0; a < A; ++a)
        {
            if (t_par) t_par[a] = 1u;
            if (o_par) o_par[a] = 1u;
            if (!counts) continue;
            for (int s = 0; s < S; ++s)
                for (int s2 = 0; s2 < S; ++s2) counts[k++] = m.phi(s, a, s2);
            for (int s2 = 0; s2 < S; ++s2)
                for (int o = 0; o < O; ++o) counts[k++] = m.psi(a, s2, o);
        }
        return;
    }

    auto m        = static_cast<FBAPOMDPState const*>(p)->model();
    auto const FS = (int)h->feat_s.size(), FO = (int)h->feat_o.size();
    IndexAction action(0);
    for (int a = 0; a < A; ++a)
    {
        action.index(a);
        for (int f = 0; f < FS; ++f)
        {
            auto const& node = m->transitionNode(&action, f);
            if (t_par) t_par[a * FS + f] = maskOf(*node.parents());
            if (counts)
                for (auto v : node._cpts) counts[k++] = v;
        }
        for (int g = 0; g < FO; ++g)
        {
            auto const& node = m->observationNode(&action, g);
            if (o_par) o_par[a * FO + g] = maskOf(*node.parents());
            if (counts)
                for (auto v : node._cpts) counts[k++] = v;
        }
    }
}

/**** MHNIPS2018 (src/beliefs/bayes-adaptive/factored/MHNIPS2018.cpp) ****/
// a fresh MHNIPS2018 belief of n particles; threshold < 0 is the log-likelihood below which
// updateEstimation runs MH. A threshold of -1e300 never triggers it (then ref_mh_run does).
int ref_mh_init(void* hv, long n, double threshold)
{
    auto h = static_cast<Handle*>(hv);
    try
    {
        if (h->mh_belief) h->mh_belief->free(*h->sim);
        h->mh_belief.reset(new RefMH((size_t)n, threshold));
        h->mh_belief->initiate(*h->sim);
    } catch (std::string const& e)
    {
        h->err = e;
        return 1;
    } catch (char const* e)
    {
        h->err = e;
        return 1;
    }
    return 0;
}

// the private MHNIPS2018::MH (MHNIPS2018.cpp:188-255) on the belief as it is
void ref_mh_run(void* hv)
{
    auto h = static_cast<Handle*>(hv);
    h->mh_belief->MH(*h->sim);
}

double ref_mh_log_likelihood(void* hv)
{
    return static_cast<Handle*>(hv)->mh_belief->_log_likelihood;
}

// FBAPOMDPPrior::computePriorModel(structure) (FBAPOMDPPrior.hpp:32 and the domain priors), the
// structure given as this repo's parent bitmasks; counts in this repo's layout. Returns the number of
// cells, -1 on error.
long ref_prior_model(void* hv, uint32_t const* t_par, uint32_t const* o_par, float* counts)
{
    auto h = static_cast<Handle*>(hv);
    try
    {
        auto const& fba = dynamic_cast<FactoredPOMDP const&>(*h->sim);
        int const A = h->sim->domainSize()->_A, FS = (int)h->feat_s.size(), FO = (int)h->feat_o.size();
        ::bayes_adaptive::factored::BABNModel::Structure st;
        st.T.resize(A), st.O.resize(A);
        for (int a = 0; a < A; ++a)
        {
            for (int f = 0; f < FS; ++f)
            {
                std::vector<int> par;
                for (int p = 0; p < FS; ++p)
                    if (t_par[a * FS + f] & (1u << p)) par.push_back(p);
                st.T[a].push_back(par);
            }
            for (int g = 0; g < FO; ++g)
            {
                std::vector<int> par;
                for (int p = 0; p < FS; ++p)
                    if (o_par[a * FO + g] & (1u << p)) par.push_back(p);
                st.O[a].push_back(par);
            }
        }
        auto model = fba.prior()->computePriorModel(st);
        long k     = 0;
        IndexAction action(0);
        for (int a = 0; a < A; ++a)
        {
            action.index(a);
            for (int f = 0; f < FS; ++f)
                for (auto v : model.transitionNode(&action, f)._cpts) counts[k++] = v;
            for (int g = 0; g < FO; ++g)
                for (auto v : model.observationNode(&action, g)._cpts) counts[k++] = v;
        }
        return k;
    } catch (std::string const& e)
    {
        h->err = e;
    } catch (char const* e)
    {
        h->err = e;
    }
    return -1;
}

// FBAPOMDP::mutate (the domain prior's mutate, FBAPOMDP.cpp:57-61) on a structure given as parent bitmasks
int ref_mutate(void* hv, uint32_t const* t_par, uint32_t const* o_par, uint32_t* t_out, uint32_t* o_out)
{
    auto h = static_cast<Handle*>(hv);
    try
    {
        auto const& fba = dynamic_cast<FactoredPOMDP const&>(*h->sim);
        int const A = h->sim->domainSize()->_A, FS = (int)h->feat_s.size(), FO = (int)h->feat_o.size();
        ::bayes_adaptive::factored::BABNModel::Structure st;
        st.T.resize(A), st.O.resize(A);
        for (int a = 0; a < A; ++a)
        {
            for (int f = 0; f < FS; ++f)
            {
                std::vector<int> par;
                for (int p = 0; p < FS; ++p)
                    if (t_par[a * FS + f] & (1u << p)) par.push_back(p);
                st.T[a].push_back(par);
            }
            for (int g = 0; g < FO; ++g)
            {
                std::vector<int> par;
                for (int p = 0; p < FS; ++p)
                    if (o_par[a * FO + g] & (1u << p)) par.push_back(p);
                st.O[a].push_back(par);
            }
        }
        auto out = fba.mutate(st);
        for (int a = 0; a < A; ++a)
        {
            for (int f = 0; f < FS; ++f) t_out[a * FS + f] = maskOf(out.T[a][f]);
            for (int g = 0; g < FO; ++g) o_out[a * FO + g] = maskOf(out.O[a][g]);
        }
    } catch (std::string const& e)
    {
        h->err = e;
        return 1;
    } catch (char const* e)
    {
        h->err = e;
        return 1;
    }
    return 0;
}

// The steps of computePosterior (MHNIPS2018.cpp:41-109) driven from here through the reference's own
// model API (BABNModel::sampleStateIndex / sampleObservationIndex / incrementCountsOf,
// FBAPOMDP::sampleDomainState) on computePriorModel(structure): a second opinion on the oracle's
// orc_mh_replay_history that shares the reference's arithmetic and RNG. Returns episode attempts.
long ref_mh_posterior_probe(void* hv, uint32_t const* t_par, uint32_t const* o_par, int n_episodes,
                            int const* episode_len, int const* actions, int const* observations, float* counts,
                            int* last_state)
{
    auto h = static_cast<Handle*>(hv);
    auto const& fba = dynamic_cast<FactoredPOMDP const&>(*h->sim);
    int const A = h->sim->domainSize()->_A, FS = (int)h->feat_s.size(), FO = (int)h->feat_o.size();
    ::bayes_adaptive::factored::BABNModel::Structure st;
    st.T.resize(A), st.O.resize(A);
    for (int a = 0; a < A; ++a)
    {
        for (int f = 0; f < FS; ++f)
        {
            std::vector<int> par;
            for (int p = 0; p < FS; ++p)
                if (t_par[a * FS + f] & (1u << p)) par.push_back(p);
            st.T[a].push_back(par);
        }
        for (int g = 0; g < FO; ++g)
        {
            std::vector<int> par;
            for (int p = 0; p < FS; ++p)
                if (o_par[a * FO + g] & (1u << p)) par.push_back(p);
            st.O[a].push_back(par);
        }
    }
    auto model  = fba.prior()->computePriorModel(st);
    auto method = rnd::sample::Dir::sampleFromExpectedMult;
    IndexState s(0), new_s(0);
    IndexAction act(0);
    IndexObservation o(0), real_o(0);
    long attempts = 0;
    int first     = 0;
    for (int e = 0; e < n_episodes; ++e)
    {
        for (;;)
        {
            ++attempts;
            auto sampled = fba.sampleDomainState();
            s.index(sampled->index());
            fba.releaseDomainState(sampled);
            std::vector<std::pair<int, int>> done;
            for (int t = 0; t < episode_len[e]; ++t)
            {
                act.index(actions[first + t]);
                new_s.index(model.sampleStateIndex(&s, &act, method));
                o.index(model.sampleObservationIndex(&act, &new_s, method));
                if (o.index() != observations[first + t]) break;
                model.incrementCountsOf(&s, &act, &o, &new_s);
                done.emplace_back(s.index(), new_s.index());
                s.index(new_s.index());
            }
            if ((int)done.size() == episode_len[e]) break;
            for (size_t t = 0; t < done.size(); ++t)
            {
                s.index(done[t].first), new_s.index(done[t].second);
                act.index(actions[first + (int)t]);
                real_o.index(observations[first + (int)t]);
                model.incrementCountsOf(&s, &act, &real_o, &new_s, -1);
            }
        }
        first += episode_len[e];
    }
    *last_state = new_s.index();
    long k = 0;
    IndexAction action(0);
    for (int a = 0; a < A; ++a)
    {
        action.index(a);
        for (int f = 0; f < FS; ++f)
            for (auto v : model.transitionNode(&action, f)._cpts) counts[k++] = v;
        for (int g = 0; g < FO; ++g)
            for (auto v : model.observationNode(&action, g)._cpts) counts[k++] = v;
    }
    return attempts;
}

/**** the composite structure beliefs ****/
// kind: F_CHEAT = CheatingReinvigoration(n, amount, threshold < 0), F_INC = StructureIncubatorSampling(n,
// amount, 0 < threshold <= 1), F_GIBBS = MHwithinGibbs(n, threshold < 0, amount: 0 = MSG, 1 = RS),
// F_NESTED = NestedBelief(n, amount)
int ref_composite_init(void* hv, int kind, long n, long amount, double threshold)
{
    auto h = static_cast<Handle*>(hv);
    try
    {
        if (kind == F_CHEAT)
        {
            if (h->cheat_belief) h->cheat_belief->free(*h->sim);
            h->cheat_belief.reset(new RefCheat((size_t)n, (size_t)amount, threshold));
            h->cheat_belief->initiate(*h->sim);
        } else if (kind == F_INC)
        {
            if (h->incub_belief) h->incub_belief->free(*h->sim);
            h->incub_belief.reset(new RefIncub((size_t)n, (size_t)amount, threshold));
            h->incub_belief->initiate(*h->sim);
        } else if (kind == F_GIBBS)
        {
            if (h->gibbs_belief) h->gibbs_belief->free(*h->sim);
            h->gibbs_belief.reset(new RefGibbs((size_t)n, threshold, amount ? RefGibbs::RS : RefGibbs::MSG));
            h->gibbs_belief->initiate(*h->sim);
        } else if (kind == F_NESTED)
        {
            if (h->nested_belief) h->nested_belief->free(*h->sim);
            h->nested_belief.reset(new RefNested((size_t)n, (size_t)amount));
            h->nested_belief->initiate(*h->sim);
        } else
        {
            h->err = "ref_composite_init: unknown kind";
            return 1;
        }
    } catch (std::string const& e)
    {
        h->err = e;
        return 1;
    } catch (char const* e)
    {
        h->err = e;
        return 1;
    }
    return 0;
}

beliefs::BABelief* compositeOf(Handle* h, int kind)
{
    switch (kind)
    {
        case F_CHEAT: return h->cheat_belief.get();
        case F_INC: return h->incub_belief.get();
        case F_GIBBS: return h->gibbs_belief.get();
        case F_NESTED: return h->nested_belief.get();
    }
    return nullptr;
}

void ref_composite_update(void* hv, int kind, int a, int o)
{
    auto h = static_cast<Handle*>(hv);
    // MHwithinGibbs keeps the pointers in its history (MHwithinGibbs.cpp:317): heap objects, released by free()
    if (kind == F_GIBBS) h->gibbs_belief->updateEstimation(new IndexAction(a), new IndexObservation(o), *h->sim);
    else
    {
        IndexAction act(a);
        IndexObservation obs(o);
        compositeOf(h, kind)->updateEstimation(&act, &obs, *h->sim);
    }
}

void ref_composite_reset(void* hv, int kind)
{
    auto h = static_cast<Handle*>(hv);
    compositeOf(h, kind)->resetDomainStateDistribution(*h->sim);
}

// weights and _total_weight of a weighted filter (F_IS, F_MH, F_CHEAT, F_INC_SHADOW, F_GIBBS, F_NESTED)
void ref_filter_weights(void* hv, int filter, double* w, double* total)
{
    auto h = static_cast<Handle*>(hv);
    auto n = filterSize(h, filter);
#define FBA_W(F)                                              \
    {                                                         \
        for (long i = 0; i < n; ++i) w[i] = (F).particle(i)->w; \
        *total = (F)._total_weight;                           \
    }
    switch (filter)
    {
        case F_IS: FBA_W(h->is_belief->_filter) break;
        case F_MH: FBA_W(h->mh_belief->_belief) break;
        case F_CHEAT: FBA_W(h->cheat_belief->_belief) break;
        case F_INC_SHADOW: FBA_W(h->incub_belief->_shadow_belief) break;
        case F_GIBBS: FBA_W(h->gibbs_belief->_belief) break;
        case F_NESTED: FBA_W(h->nested_belief->_filter) break;
        default: *total = -1;
    }
#undef FBA_W
}

// NestedBelief: the bottom filters' domain states, [n_top][n_bottom] row-major
void ref_nested_states(void* hv, int* out)
{
    auto h  = static_cast<Handle*>(hv);
    auto& f = h->nested_belief->_filter;
    long k  = 0;
    for (size_t i = 0; i < f.size(); ++i)
        for (auto s : f.particle(i)->particle.second.particles()) out[k++] = s->index();
}

// NestedBelief::sample (NestedBelief.cpp:117-127): out[0] = index of the drawn top particle, out[1] = its domain state
void ref_nested_sample(void* hv, int* out)
{
    auto h   = static_cast<Handle*>(hv);
    auto s   = static_cast<BAState const*>(h->nested_belief->sample());
    auto& f  = h->nested_belief->_filter;
    out[0]   = -1;
    for (size_t i = 0; i < f.size(); ++i)
        if (f.particle(i)->particle.first == s) out[0] = (int)i;
    out[1] = s->_domain_state->index();
}

// the private MHwithinGibbs::reinvigorate (MHwithinGibbs.cpp:334-395) on the belief as it is
void ref_gibbs_run(void* hv)
{
    auto h = static_cast<Handle*>(hv);
    h->gibbs_belief->reinvigorate(*h->sim);
}

double ref_gibbs_log_likelihood(void* hv)
{
    return static_cast<Handle*>(hv)->gibbs_belief->_log_likelihood;
}

// FBAPOMDP::domainStatePrior()->prob(s) for every domain state (the prior of s_0 in msgSampleStateHistory)
void ref_state_prior(void* hv, float* out)
{
    auto h          = static_cast<Handle*>(hv);
    auto const& fba = dynamic_cast<FactoredPOMDP const&>(*h->sim);
    auto p          = fba.domainStatePrior();
    for (int s = 0; s < h->sim->domainSize()->_S; ++s) out[s] = p->prob(s);
}

double ref_cheat_likelihood(void* hv)
{
    return static_cast<Handle*>(hv)->cheat_belief->_likelihood;
}

// the parts of StructureIncubatorSampling::updateEstimation one by one (StructureIncubatorSampling.cpp:107-137)
// part 0: reinvigorateBelief, 1: reinvigorateShadowBelief, 2: rejectSample(_belief), 3: rejectSample(_fully_connected),
// 4: importance_sampling::update(_shadow), 5: importance_sampling::resample(_shadow). Returns part 4's likelihood.
double ref_incubator_part(void* hv, int part, int a, int o)
{
    auto h        = static_cast<Handle*>(hv);
    auto& b       = *h->incub_belief;
    auto const& f = dynamic_cast<FactoredPOMDP const&>(*h->sim);
    IndexAction act(a);
    IndexObservation obs(o);
    switch (part)
    {
        case 0: b.reinvigorateBelief(f); break;
        case 1: b.reinvigorateShadowBelief(f); break;
        case 2: ::beliefs::rejectSample(&act, &obs, *h->sim, b._size, b._belief); break;
        case 3: ::beliefs::rejectSample(&act, &obs, *h->sim, b._size, b._fully_connected_belief); break;
        case 4: return ::beliefs::importance_sampling::update(b._shadow_belief, &act, &obs, *h->sim);
        case 5: ::beliefs::importance_sampling::resample(b._shadow_belief, *h->sim, b._size); break;
    }
    return 0.0;
}

// overwrite the weights of the incubator's shadow belief (to exercise promotion and leastLikely on
// non-uniform weights, which the reference's own update never produces: it resamples every step)
void ref_incubator_set_shadow_weights(void* hv, double const* w)
{
    auto h  = static_cast<Handle*>(hv);
    auto& f = h->incub_belief->_shadow_belief;
    double total = 0;
    for (size_t i = 0; i < f.size(); ++i)
    {
        f.particle(i)->w = w[i];
        total += w[i];
    }
    f._total_weight = total;
}

void ref_cheat_only(void* hv)
{
    auto h = static_cast<Handle*>(hv);
    h->cheat_belief->cheat(*h->sim);
}

/**** importance sampling ****/
// importance_sampling::update only (weights stay un-resampled); returns the step likelihood
double ref_is_update(void* hv, int a, int o)
{
    auto h = static_cast<Handle*>(hv);
    IndexAction act(a);
    IndexObservation obs(o);
    return beliefs::importance_sampling::update(h->is_belief->_filter, &act, &obs, *h->sim);
}

void ref_is_resample(void* hv)
{
    auto h = static_cast<Handle*>(hv);
    beliefs::importance_sampling::resample(h->is_belief->_filter, *h->sim, h->is_belief->_n);
}

// the full Belief::updateEstimation of the given belief kind
void ref_update_estimation(void* hv, int kind, int a, int o)
{
    auto h = static_cast<Handle*>(hv);
    IndexAction act(a);
    IndexObservation obs(o);
    if (kind == F_IS) h->is_belief->updateEstimation(&act, &obs, *h->sim);
    else if (kind == F_RS)
        h->rs_belief->updateEstimation(&act, &obs, *h->sim);
    else if (kind == F_MH)
    { // MHNIPS2018 keeps the pointers in its history (MHNIPS2018.cpp:173) and the domains' copyAction /
      // copyObservation hand the same pointer back (e.g. FactoredTiger.cpp:49-54,136-141): they must outlive
      // this call. The belief's free() releases them through the domain.
        h->mh_belief->updateEstimation(new IndexAction(a), new IndexObservation(o), *h->sim);
    }
    else
        h->reinv_belief->updateEstimation(&act, &obs, *h->sim);
}

void ref_reset_domain_states(void* hv, int kind)
{
    auto h = static_cast<Handle*>(hv);
    if (kind == F_IS) h->is_belief->resetDomainStateDistribution(*h->sim);
    else if (kind == F_RS)
        h->rs_belief->resetDomainStateDistribution(*h->sim);
    else if (kind == F_MH)
        h->mh_belief->resetDomainStateDistribution(*h->sim);
    else
        h->reinv_belief->resetDomainStateDistribution(*h->sim);
}

// reinvigorateParticles alone (ReinvigoratingRejectionSampling.cpp:121-131)
void ref_reinvigorate_only(void* hv)
{
    auto h = static_cast<Handle*>(hv);
    h->reinv_belief->reinvigorateParticles(*h->sim);
}

// the two rejectSample calls of ReinvigoratingRejectionSampling::updateEstimation (:100-101)
void ref_reinv_reject_only(void* hv, int a, int o)
{
    auto h = static_cast<Handle*>(hv);
    IndexAction act(a);
    IndexObservation obs(o);
    auto& b = *h->reinv_belief;
    ::beliefs::rejectSample(&act, &obs, *h->sim, b._size, b._belief);
    ::beliefs::rejectSample(&act, &obs, *h->sim, b._size, b._fully_connected_belief);
}

/**** one hyper-state step on particle i, in place (BAPOMDP::step, BAPOMDP.cpp:111-143) ****/
// mode: 0 = UpdateCounts, 1 = KeepCounts. out[0] = new state, out[1] = observation, out[2] = terminal
double ref_particle_step(void* hv, int filter, long i, int a, int mode, int* out)
{
    auto h = static_cast<Handle*>(hv);
    IndexAction act(a);
    State const* s = particleOf(h, filter, i);
    Observation const* o(nullptr);
    Reward r(0);
    auto t = h->sim->step(
        &s, &act, &o, &r, mode ? BAPOMDP::StepType::KeepCounts : BAPOMDP::StepType::UpdateCounts);
    out[0] = static_cast<BAState const*>(s)->_domain_state->index();
    out[1] = o->index();
    out[2] = t.terminated();
    h->sim->releaseObservation(o);
    return r.toDouble();
}

double ref_particle_obs_prob(void* hv, int filter, long i, int a, int o)
{
    auto h = static_cast<Handle*>(hv);
    IndexAction act(a);
    IndexObservation obs(o);
    return h->sim->computeObservationProbability(&obs, &act, particleOf(h, filter, i));
}

/**** BABNModel::LogBDScore (BABNModel.cpp:451-478) of particle i of `filter` against particle j of
      `prior_filter` (same structure required; factored models only) ****/
double ref_log_bd_score(void* hv, int filter, long i, int prior_filter, long j)
{
    auto h = static_cast<Handle*>(hv);
    if (!h->factored) return 0.0;
    auto a = const_cast<FBAPOMDPState*>(static_cast<FBAPOMDPState const*>(particleOf(h, filter, i)))->model();
    auto b = const_cast<FBAPOMDPState*>(static_cast<FBAPOMDPState const*>(particleOf(h, prior_filter, j)))->model();
    return a->LogBDScore(*b);
}

/**** rollouts (RBAPOUCT::rollout, RBAPOUCT.cpp:295-323) ****/
// Mirrors what selectAction does around a simulation (RBAPOUCT.cpp:86-106): KeepCounts, the
// particle is not copied, its domain state is set to start_state and restored afterwards.
double ref_rollout(void* hv, int filter, long i, int start_state, int depth)
{
    auto h = static_cast<Handle*>(hv);
    if (!h->planner) h->planner.reset(new planners::RBAPOUCT(h->conf));

    auto particle  = particleOf(h, filter, i);
    auto old_mode  = h->sim->mode();
    auto old_state = particle->_domain_state;
    h->sim->mode(BAPOMDP::StepType::KeepCounts);
    const_cast<BAState*>(particle)->_domain_state =
        h->sim->copyDomainState(h->sim->domainState(start_state));

    auto ret = h->planner->rollout(particle, *h->sim, depth);

    h->sim->releaseDomainState(particle->_domain_state);
    const_cast<BAState*>(particle)->_domain_state = old_state;
    h->sim->mode(old_mode);
    return ret.toDouble();
}

/**** domain functors (BADomainExtension::{reward,terminal}) ****/
double ref_reward(void* hv, int s, int a, int s2, int* terminal)
{
    auto h = static_cast<Handle*>(hv);
    IndexAction act(a);
    auto ext  = h->sim->_ba_domain_ext.get();
    auto st   = ext->getState(s);
    auto st2  = ext->getState(s2);
    *terminal = ext->terminal(st, &act, st2).terminated();
    return ext->reward(st, &act, st2).toDouble();
}

/**** domain description in this repo's vocabulary (see oracle/fba_oracle.h ORC_DOM_* etc.) ****/
// out_i: [0] domain kind, [1] action draw kind, [2] start kind, [3..6] start_ip[4], [8..39] dom_ip[32]
// out_d: [0..7] dom_dp, [8] start_total. start_values: float[S] (categorical), start_table: int[64]
void ref_domain_desc(void* hv, int* out_i, double* out_d, float* start_values, int* start_table)
{
    auto h        = static_cast<Handle*>(hv);
    auto const& d = h->conf.domain_conf.domain;
    auto const S  = h->sim->domainSize()->_S;
    int* ip       = out_i + 8;
    for (int i = 0; i < 40; ++i) out_i[i] = 0;
    for (int i = 0; i < 9; ++i) out_d[i] = 0;

    if (d == "episodic-tiger" || d == "continuous-tiger")
    {
        out_i[0] = 1;
        ip[0]    = (d == "episodic-tiger");
        out_i[2] = 1; // boolean() ? LEFT : RIGHT (Tiger.cpp:18)
        out_i[3] = 0;
        out_i[4] = 1;
    } else if (d == "episodic-factored-tiger" || d == "continuous-factored-tiger")
    {
        out_i[0] = 2;
        ip[0]    = (d == "episodic-factored-tiger");
        out_i[2] = 2; // uniform_int over S (FactoredTiger.cpp:74)
        out_i[3] = S;
    } else if (d == "linear-sysadmin" || d == "independent-sysadmin")
    {
        out_i[0] = 3;
        ip[0]    = (int)h->conf.domain_conf.size;
        out_d[0] = domains::SysAdmin::param._reboot_cost;
        out_i[2] = 0; // constant: all computers up (SysAdmin.cpp:102-105)
        out_i[3] = S - 1;
    } else if (d == "gridworld")
    {
        out_i[0]   = 4;
        out_i[1]   = 1; // slowRandomInt (GridWorld.cpp:223)
        auto size  = (int)h->conf.domain_conf.size;
        auto goals = domains::GridWorld::goalLocations(size);
        ip[0]      = size;
        ip[1]      = (int)goals.size();
        for (size_t g = 0; g < goals.size(); ++g)
        {
            ip[2 + 2 * g] = goals[g].x;
            ip[3 + 2 * g] = goals[g].y;
        }
        out_d[0] = domains::GridWorld::goal_reward;
        out_d[1] = domains::GridWorld::step_reward;
        out_i[2] = 3; // two slowRandomInt draws (GridWorld.cpp:265-267)
        out_i[3] = (int)domains::GridWorld::start_locations.size();
        out_i[4] = (int)goals.size();
        auto gw  = dynamic_cast<domains::GridWorld const*>(h->env.get());
        int k    = 0;
        for (auto const& sl : domains::GridWorld::start_locations)
            for (auto const& g : goals) start_table[k++] = gw->getState(sl, g)->index();
    } else if (d == "centered-collision-avoidance" || d == "random-collision-avoidance")
    {
        out_i[0] = 5;
        ip[0]    = (int)h->conf.domain_conf.width;
        ip[1]    = (int)h->conf.domain_conf.height;
        ip[2]    = (int)h->conf.domain_conf.size;
        out_d[0] = domains::CollisionAvoidance::MOVE_PENALTY;
        out_d[1] = domains::CollisionAvoidance::COLLIDE_PENALTY;
        out_i[2] = 4; // categorical over S (CollisionAvoidance.cpp:274)
        out_i[3] = S;
        auto ca  = dynamic_cast<domains::CollisionAvoidance const*>(h->env.get());
        for (int i = 0; i < S; ++i) start_values[i] = ca->_state_prior._values[i];
        out_d[8] = ca->_state_prior._total;
    } else
    {
        out_i[0] = -1;
    }
}

/**** (a,o) script from the TRUE environment under a uniformly random policy ****/
// Runs episodes of at most `horizon` steps until `steps` entries are filled.
// flags[t] bit0 = terminal after step t, bit1 = first step of an episode.
void ref_env_script(void* hv, int steps, int horizon, int* actions, int* observations, int* flags)
{
    auto h   = static_cast<Handle*>(hv);
    auto env = h->env.get();
    auto dom = dynamic_cast<POMDP const*>(env); // every domain here is a POMDP (test/test.cpp:72)
    int t    = 0;
    while (t < steps)
    {
        State const* s = env->sampleStartState();
        bool terminal  = false;
        for (int k = 0; k < horizon && !terminal && t < steps; ++k, ++t)
        {
            auto a = dom->generateRandomAction(s);
            Observation const* o(nullptr);
            Reward r(0);
            terminal        = env->step(&s, a, &o, &r).terminated();
            actions[t]      = a->index();
            observations[t] = o->index();
            flags[t]        = (terminal ? 1 : 0) | (k == 0 ? 2 : 0);
            dom->releaseAction(a);
            env->releaseObservation(o);
        }
        env->releaseState(s);
    }
}


/**** drop-in check: the reference's own episode loop + planner, with this repo's CUDA belief ****/
// Runs `episodes` episodes of experiment::bapomdp::run's inner loop (BAPOMDPExperiment.cpp:46-75):
// initiate once, then per episode resetDomainStateDistribution + episode::run with the reference's
// planner (--planner string). kind: 0 = the reference's BAImportanceSampling (CPU),
// 1 = fba_b200::CudaBAImportanceSampling, 2 = the reference's BARejectionSampling,
// 3 = fba_b200::CudaBARejectionSampling, 4 = the reference's ReinvigoratingRejectionSampling,
// 5 = fba_b200::CudaReinvigoratingRejectionSampling (amount = n / 8, mutate kind from the domain).
// returns[e] = discounted return of episode e.
// rc: 0 ok, 1 error (see ref_error).
int ref_adapter_episodes(void* hv, int kind, long n, char const* planner, int sims, int episodes,
                         double* returns)
{
    auto h = static_cast<Handle*>(hv);
    try
    {
        auto conf                                = h->conf;
        conf.planner                             = planner;
        conf.planner_conf.mcts_simulation_amount = sims;
        conf.planner_conf.mcts_max_depth         = conf.horizon;
        // planner "cuda-po-uct[:wave]" = this repo's wave-parallel POMCP over the C ABI (tree on the host),
        // "cuda-tree-po-uct[:wave]" = the same with the tree on the device
        std::unique_ptr<Planner> plan;
        if (conf.planner.rfind("cuda-po-uct", 0) == 0 || conf.planner.rfind("cuda-tree-po-uct", 0) == 0)
        {
            int wave = 64;
            auto pos = conf.planner.find(':');
            if (pos != std::string::npos) wave = std::stoi(conf.planner.substr(pos + 1));
            if (conf.planner.rfind("cuda-tree", 0) == 0) plan.reset(new fba_b200::CudaTreePOUCT(conf, wave));
            else
                plan.reset(new fba_b200::CudaBatchedPOUCT(conf, wave));
        } else
            plan = factory::makeBAPlanner(conf);

        std::unique_ptr<beliefs::BABelief> belief;
        if (kind == 0) belief.reset(new beliefs::BAImportanceSampling(n));
        else if (kind == 1)
            belief.reset(new fba_b200::CudaBAImportanceSampling(n));
        else if (kind == 2)
            belief.reset(new beliefs::BARejectionSampling(n));
        else if (kind == 3)
            belief.reset(new fba_b200::CudaBARejectionSampling(n));
        else if (kind == 6) // the reference's MHNIPS2018; threshold chosen so that MH runs every few steps
            belief.reset(new RefMH((size_t)n, -4.0));
        else if (kind == 7)
            belief.reset(new fba_b200::CudaMHNIPS2018((size_t)n, -4.0));
        else
        {
            auto const& d  = h->conf.domain_conf.domain;
            int const mut  = (d.find("factored-tiger") != std::string::npos)       ? FBA_MUT_FACTORED_TIGER
                             : (d.find("collision-avoidance") != std::string::npos) ? FBA_MUT_COLLISION_AVOIDANCE
                             : (d.find("sysadmin") != std::string::npos)            ? FBA_MUT_SYSADMIN
                                                                                    : FBA_MUT_GRIDWORLD;
            size_t const k = std::max<size_t>(1, n / 8);
            if (kind == 4) belief.reset(new ReinvRS(n, k));
            else if (kind == 5)
                belief.reset(new fba_b200::CudaReinvigoratingRejectionSampling(n, k, mut));
            // the composite structure beliefs with the settings of the reference's own integration tests
            // (test/test.cpp:296-353: cheating threshold -3, incubator threshold .05)
            else if (kind == 8)
                belief.reset(new RefCheat((size_t)n, k, -3.0));
            else if (kind == 9)
                belief.reset(new fba_b200::CudaCheatingReinvigoration((size_t)n, k, -3.0));
            else if (kind == 10)
                belief.reset(new RefIncub((size_t)n, k, 0.05));
            else if (kind == 11)
                belief.reset(new fba_b200::CudaStructureIncubatorSampling((size_t)n, k, 0.05, mut));
            // NestedBelief sized as the factory does (BABelief.cpp:66-69): n top particles, n * n bottom ones each
            else if (kind == 12)
                belief.reset(new RefNested((size_t)n, (size_t)(n * n)));
            else if (kind == 13)
                belief.reset(new fba_b200::CudaNestedBelief((size_t)n, (size_t)(n * n)));
            // MHwithinGibbs, threshold -4 so that the chain runs every few steps; 14/15 message passing, 16/17 rejection
            else if (kind == 14 || kind == 16)
                belief.reset(new RefGibbs((size_t)n, -4.0, kind == 14 ? RefGibbs::MSG : RefGibbs::RS));
            else if (kind == 15 || kind == 17)
                belief.reset(new fba_b200::CudaMHwithinGibbs(
                    (size_t)n, -4.0, kind == 15 ? fba_b200::CudaMHwithinGibbs::MSG : fba_b200::CudaMHwithinGibbs::RS));
            else
                throw std::string("ref_adapter_episodes: unknown belief kind");
        }

        belief->initiate(*h->sim);
        h->update_seconds = 0, h->update_calls = 0;
        h->update_times.clear();
        TimedBelief timed(belief.get(), h);
        // the CUDA planners need to see the CUDA belief itself (they read its device handle): no timing wrapper there
        bool const cuda_planner = conf.planner.rfind("cuda-", 0) == 0;
        beliefs::BABelief& used = cuda_planner ? *belief : static_cast<beliefs::BABelief&>(timed);
        for (int e = 0; e < episodes; ++e)
        {
            used.resetDomainStateDistribution(*h->sim);
            auto r = episode::run(
                *plan, used, *h->env, *h->sim, Horizon(conf.horizon), Discount(conf.discount));
            returns[e] = r.ret.toDouble();
        }
        // how often the structure-learning part of a CUDA belief ran (MH runs, Gibbs chains, cheats)
        h->adapter_events = -1;
        if (auto b = dynamic_cast<fba_b200::CudaMHNIPS2018*>(belief.get())) h->adapter_events = (long)b->mhRuns();
        if (auto b = dynamic_cast<fba_b200::CudaMHwithinGibbs*>(belief.get())) h->adapter_events = (long)b->chains();
        if (auto b = dynamic_cast<fba_b200::CudaCheatingReinvigoration*>(belief.get())) h->adapter_events = (long)b->cheats();
        belief->free(*h->sim);
    } catch (std::string const& e)
    {
        h->err = e;
        return 1;
    } catch (char const* e)
    {
        h->err = e;
        return 1;
    }
    return 0;
}

long ref_adapter_events(void* hv)
{
    return static_cast<Handle*>(hv)->adapter_events;
}

// per-call seconds inside Belief::updateEstimation during the last ref_adapter_episodes (at most cap of them)
long ref_adapter_update_times(void* hv, double* out, long cap)
{
    auto h = static_cast<Handle*>(hv);
    long n = std::min<long>(cap, (long)h->update_times.size());
    for (long i = 0; i < n; ++i) out[i] = h->update_times[(size_t)i];
    return (long)h->update_times.size();
}

// seconds spent inside Belief::updateEstimation during the last ref_adapter_episodes, and the number of calls
double ref_adapter_update_seconds(void* hv, long* calls)
{
    auto h = static_cast<Handle*>(hv);
    *calls = h->update_calls;
    return h->update_seconds;
}


// fba_b200::runBatchedExperiment (host/CudaExperiment.hpp): `runs` runs of `episodes` episodes in
// lockstep on the GPU, n particles and `sims` simulations each. returns: episodes x runs, row-major.
// Result: seconds of wall time, < 0 on error.
double ref_batched_episodes_on(void* hv, long n, int runs, int sims, int episodes, int sims_per_wave, int device,
                               unsigned long long seed, double* returns);
double ref_batched_episodes(void* hv, long n, int runs, int sims, int episodes, int sims_per_wave, double* returns)
{
    return ref_batched_episodes_on(hv, n, runs, sims, episodes, sims_per_wave, 0, 4711ull, returns);
}

// the same on a given GPU with a given seed: independent runs shard over GPUs with no communication
double ref_batched_episodes_on(void* hv, long n, int runs, int sims, int episodes, int sims_per_wave, int device,
                               unsigned long long seed, double* returns)
{
    auto h = static_cast<Handle*>(hv);
    try
    {
        auto conf                                = h->conf;
        conf.belief_conf.particle_amount         = n;
        conf.planner_conf.mcts_simulation_amount = sims;
        conf.planner_conf.mcts_max_depth         = conf.horizon;
        conf.num_episodes                        = episodes;
        auto t0  = std::chrono::steady_clock::now();
        auto res = fba_b200::runBatchedExperiment(*h->sim, conf, runs, sims_per_wave, seed, device);
        double const dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (int e = 0; e < episodes; ++e)
            for (int r = 0; r < runs; ++r) returns[(size_t)e * runs + r] = res[e][r];
        return dt;
    } catch (std::string const& e)
    {
        h->err = e;
    } catch (char const* e)
    {
        h->err = e;
    } catch (std::exception const& e)
    {
        h->err = e.what();
    }
    return -1.0;
}

// the same, writing the reference's result file (experiment::bapomdp::Result::log, BAPOMDPExperiment.cpp:20-30,
// what bapomdp.cpp:43-44 writes to --output-file) to `path`
double ref_batched_experiment_file(void* hv, long n, int runs, int sims, int episodes, unsigned long long seed,
                                   char const* path, double* returns)
{
    auto h = static_cast<Handle*>(hv);
    try
    {
        auto conf                                = h->conf;
        conf.belief_conf.particle_amount         = n;
        conf.planner_conf.mcts_simulation_amount = sims;
        conf.planner_conf.mcts_max_depth         = conf.horizon;
        conf.num_episodes                        = episodes;
        experiment::bapomdp::Result result(episodes);
        auto t0  = std::chrono::steady_clock::now();
        auto res = fba_b200::runBatchedExperiment(*h->sim, conf, runs, 1, seed, 0, &result);
        double const dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (int e = 0; e < episodes; ++e)
            for (int r = 0; r < runs; ++r) returns[(size_t)e * runs + r] = res[e][r];
        std::ofstream f(path);
        result.log(f);
        return dt;
    } catch (std::string const& e)
    {
        h->err = e;
    } catch (char const* e)
    {
        h->err = e;
    } catch (std::exception const& e)
    {
        h->err = e.what();
    }
    return -1.0;
}

// Belief::initiate of the CUDA adapter (fba_b200::CudaBAImportanceSampling, n particles) next to n_ref
// draws of the reference's own BAPOMDP::sampleStartState: domain states and structure ids of both (the
// ids come from the adapter's own structure table, so they are comparable), the adapter's wall time and
// how many prior samples it took on the host. Returns 0 on success.
int ref_adapter_initiate(void* hv, long n, long n_ref, double* seconds, long* host_samples, int* cuda_state,
                         int* cuda_sid, int* ref_state, int* ref_sid)
{
    auto h = static_cast<Handle*>(hv);
    try
    {
        fba_b200::CudaBAImportanceSampling belief((size_t)n);
        auto const t0 = std::chrono::steady_clock::now();
        belief.initiate(*h->sim);
        *seconds      = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *host_samples = (long)belief.hostPriorSamples();
        if (fba_belief_download(belief.handle(), 0, n, cuda_state, cuda_sid, nullptr, nullptr) != FBA_OK)
            throw std::string("download failed");
        std::vector<float> block;
        for (long i = 0; i < n_ref; ++i)
        {
            auto p       = static_cast<BAState const*>(h->sim->sampleStartState());
            ref_state[i] = p->_domain_state->index();
            ref_sid[i]   = belief.cuda().describe(p, &block);
            h->sim->releaseState(p);
        }
        belief.free(*h->sim);
    } catch (std::string const& e)
    {
        h->err = e;
        return 1;
    } catch (char const* e)
    {
        h->err = e;
        return 1;
    }
    return 0;
}


// seconds per Planner::selectAction (empty history) with `sims` simulations, belief kind as in
// ref_adapter_episodes, planner "po-uct" (the reference's RBAPOUCT) or "cuda-po-uct[:wave]".
double ref_plan_seconds(void* hv, int kind, long n, char const* planner, int sims, int reps)
{
    auto h = static_cast<Handle*>(hv);
    try
    {
        auto conf                                = h->conf;
        conf.planner                             = planner;
        conf.planner_conf.mcts_simulation_amount = sims;
        conf.planner_conf.mcts_max_depth         = conf.horizon;
        std::unique_ptr<Planner> plan;
        if (conf.planner.rfind("cuda-po-uct", 0) == 0 || conf.planner.rfind("cuda-tree-po-uct", 0) == 0)
        {
            int wave = 64;
            auto pos = conf.planner.find(':');
            if (pos != std::string::npos) wave = std::stoi(conf.planner.substr(pos + 1));
            if (conf.planner.rfind("cuda-tree", 0) == 0) plan.reset(new fba_b200::CudaTreePOUCT(conf, wave));
            else
                plan.reset(new fba_b200::CudaBatchedPOUCT(conf, wave));
        } else
            plan = factory::makeBAPlanner(conf);
        std::unique_ptr<beliefs::BABelief> belief;
        if (kind == 0) belief.reset(new beliefs::BAImportanceSampling(n));
        else
            belief.reset(new fba_b200::CudaBAImportanceSampling(n));
        belief->initiate(*h->sim);
        History hist;
        auto warm = plan->selectAction(*h->sim, *belief, hist);
        h->sim->releaseAction(warm);
        auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; ++r)
        {
            auto a = plan->selectAction(*h->sim, *belief, hist);
            h->sim->releaseAction(a);
        }
        double const dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        belief->free(*h->sim);
        return dt / reps;
    } catch (std::string const& e)
    {
        h->err = e;
        return -1;
    } catch (char const* e)
    {
        h->err = e;
        return -1;
    }
}

} // extern "C"
