// C++ host adapters: the reference's own belief interfaces (samkatt/fba-pomdp) over the C ABI of
// libfba_b200.so. A maintainer of the reference adds this header to the tree, links
// libfba_b200.so, and registers the classes in factory::makeBABelief (src/beliefs/bayes-adaptive/
// BABelief.cpp:16-73) under new --belief names — see INTEGRATION.md.
//
//   CudaBAImportanceSampling : beliefs::BABelief   stands in for beliefs::BAImportanceSampling
//   CudaBARejectionSampling  : beliefs::BABelief   stands in for beliefs::BARejectionSampling
//   CudaReinvigoratingRejectionSampling : beliefs::BABelief   stands in for
//       beliefs::bayes_adaptive::factored::ReinvigoratingRejectionSampling
//
// Contract kept (src/beliefs/Belief.hpp:17-41, src/beliefs/bayes-adaptive/BABelief.hpp:22-35):
// initiate / free / sample / updateEstimation / resetDomainStateDistribution with the same
// argument meaning; errors are thrown as std::string like the reference's own beliefs do;
// sample() returns a BORROWED BAState* owned by the belief (valid until the next sample / update),
// whose _domain_state may be poked and restored by RBAPOUCT (RBAPOUCT.cpp:92-106).
//
// Everything domain specific is obtained by PROBING the reference's own objects, so any domain the
// reference supports works unchanged:
//   * reward / terminal: BADomainExtension::{reward,terminal} evaluated on all (s,a,s') once and
//     stored as the two separable tables of FBA_DOM_TABLE (an error is thrown if a domain's reward
//     is not of the form f(s,a) + g(a,s'));
//   * prior: the reference's BAPrior runs on the host (BAPOMDP::sampleStartState) and its particles
//     are uploaded — priors stay bit-identical;
//   * domain start states (initiate, resetDomainStateDistribution): drawn by the reference's own
//     domain on the host (N cheap calls) and uploaded.
// The heavy per-particle work — step, likelihood, normalisation, resampling, copies — runs on the GPU.
//
// This header only uses the reference's PUBLIC API. It is compiled and exercised by
// oracle/ref_harness.cpp (ref_adapter_episodes) so that it cannot rot.
#ifndef FBA_B200_CUDA_BELIEFS_HPP
#define FBA_B200_CUDA_BELIEFS_HPP

#include <chrono>
#include <cstdint>
#include <map>
#include <random>
#include <unordered_map>
#include <memory>
#include <string>
#include <vector>

#include "fba_pomdp_b200.h"

#include "bayes-adaptive/models/factored/FBAPOMDP.hpp"
#include "bayes-adaptive/models/table/BAPOMDP.hpp"
#include "bayes-adaptive/states/factored/BABNModel.hpp"
#include "bayes-adaptive/states/factored/FBAPOMDPState.hpp"
#include "bayes-adaptive/states/table/BAPOMDPState.hpp"
#include "beliefs/bayes-adaptive/BABelief.hpp"
#include "environment/Action.hpp"
#include "environment/Observation.hpp"
#include "environment/State.hpp"
#include "utils/index.hpp"
#include "utils/random.hpp"

namespace fba_b200 {

inline void check(fba_ctx* ctx, int rc, char const* what)
{
    if (rc != FBA_OK)
        throw std::string("fba_b200: ") + what + ": " + (ctx ? fba_last_error(ctx) : "no context");
}

// One GPU context + the uploaded description of one BAPOMDP / FBAPOMDP.
class CudaSimulator
{
public:
    // delta_capacity: -1 = choose automatically (base+delta storage for tabular models whose dense
    // count block exceeds 256 KB, with room for 2048 increments = 1024 updates per particle)
    // start_samples > 0: the domain's start-state distribution is estimated from that many draws of the
    // reference's own domain (BAPOMDP::sampleDomainState) and uploaded as a categorical, so that start
    // states can be drawn ON the device (fba_runs_reset_domain_states for thousands of runs); 0: start
    // states always come from the host domain (the single-belief adapters)
    CudaSimulator(BAPOMDP const& sim, int device = 0, int max_structures = 4096, int delta_capacity = -1,
                  int start_samples = 0) :
            _sim(sim)
    {
        if (fba_abi_version() != FBA_ABI_VERSION)
            throw std::string("fba_b200: libfba_b200.so was built with another ABI version than this header");
        int rc = fba_ctx_create(device, &_ctx);
        if (rc == FBA_ERR_NO_DEVICE) throw std::string("fba_b200: no CUDA device (there is no CPU fallback)");
        check(nullptr, rc, "fba_ctx_create");

        auto const* size = sim.domainSize();
        _fbapomdp        = dynamic_cast<::bayes_adaptive::factored::FBAPOMDP const*>(&sim);

        fba_model_desc d = {};
        d.S = size->_S, d.A = size->_A, d.O = size->_O;
        if (_fbapomdp)
        {
            _feat_s = _fbapomdp->domainFeatureSize()->_S;
            _feat_o = _fbapomdp->domainFeatureSize()->_O;
        } else
        {
            _feat_s = {size->_S};
            _feat_o = {size->_O};
        }
        d.tabular          = _fbapomdp ? 0 : 1;
        d.n_state_features = (int)_feat_s.size();
        d.n_obs_features   = (int)_feat_o.size();
        for (size_t i = 0; i < _feat_s.size(); ++i) d.state_feature_sizes[i] = _feat_s[i];
        for (size_t i = 0; i < _feat_o.size(); ++i) d.obs_feature_sizes[i] = _feat_o[i];

        probeRewards(d);
        d.domain      = FBA_DOM_TABLE;
        long long const dense_cells = (long long)d.A * d.S * d.S + (long long)d.A * d.S * d.O;
        if (delta_capacity < 0) delta_capacity = (d.tabular && dense_cells * 4 > (256 << 10)) ? 2048 : 0;
        d.delta_capacity = d.tabular ? delta_capacity : 0;
        // --dirichlet_sampling_method regular (BAPOMDP.cpp:197-201): every draw first samples the
        // multinomial from the particle's Dirichlets instead of using their expectation
        d.dirichlet_sampling = sampledDirichlets() ? 1 : 0;
        d.action_draw = FBA_ACT_UNIFORM_INT; // rollouts only; all reference domains draw uniformly
        d.start_kind  = FBA_START_CONST;     // unused: start states come from the host domain
        if (start_samples > 0)
        {
            _start_freq.assign((size_t)d.S, 0.0f);
            for (int i = 0; i < start_samples; ++i)
            {
                auto st = sim.sampleDomainState();
                _start_freq[(size_t)st->index()] += 1.0f;
                sim.releaseDomainState(st);
            }
            d.start_kind   = FBA_START_CATEGORICAL;
            d.start_ip[0]  = d.S;
            d.start_values = _start_freq.data();
            d.start_total  = (double)start_samples;
        }

        check(_ctx, fba_model_create(_ctx, &d, max_structures, &_model), "fba_model_create");
        _steps.reset(new ::bayes_adaptive::factored::BABNModel::Indexing_Steps(
            indexing::stepSize(_feat_s), indexing::stepSize(_feat_o)));
    }

    ~CudaSimulator()
    {
        fba_tree_destroy(_tree);
        fba_model_destroy(_model);
        fba_ctx_destroy(_ctx);
    }
    CudaSimulator(CudaSimulator const&) = delete;
    CudaSimulator& operator=(CudaSimulator const&) = delete;

    fba_ctx* ctx() const { return _ctx; }
    fba_model* model() const { return _model; }
    // the device-resident POMCP tree for this simulator (CudaTreePOUCT); it lives and dies with the
    // context, so a planner that outlives the belief never holds a dangling handle
    fba_tree* tree(int64_t simulations, int depth) const
    {
        if (_tree && (simulations > _tree_sims || depth > _tree_depth))
        {
            fba_tree_destroy(_tree);
            _tree = nullptr;
        }
        if (!_tree)
        {
            check(_ctx, fba_tree_create(_ctx, _model, simulations, depth, &_tree), "fba_tree_create");
            _tree_sims  = simulations;
            _tree_depth = depth;
        }
        return _tree;
    }
    BAPOMDP const& sim() const { return _sim; }
    bool factored() const { return _fbapomdp != nullptr; }
    int A() const { return _sim.domainSize()->_A; }
    int S() const { return _sim.domainSize()->_S; }
    int O() const { return _sim.domainSize()->_O; }
    int FS() const { return (int)_feat_s.size(); }
    int FO() const { return (int)_feat_o.size(); }

    // structure id + count block (this repo's layout) of one reference particle
    int32_t describe(BAState const* p, std::vector<float>* counts) const
    {
        std::vector<uint32_t> tp((size_t)A() * FS()), op((size_t)A() * FO());
        counts->clear();
        if (!factored())
        {
            // BAFlatModel::count is the only public cell accessor; it is non-const because it may
            // materialise a copy-on-write row (BAFlatModel.cpp:185-252) — values are unchanged
            auto model = const_cast<BAPOMDPState*>(static_cast<BAPOMDPState const*>(p))->model();
            IndexAction a(0);
            IndexState s(0), s2(0);
            IndexObservation o(0);
            for (int ai = 0; ai < A(); ++ai)
            {
                a.index(ai);
                tp[ai] = 1u, op[ai] = 1u;
                for (int si = 0; si < S(); ++si)
                    for (int ti = 0; ti < S(); ++ti)
                    {
                        s.index(si), s2.index(ti);
                        counts->push_back(model->count(&s, &a, &s2));
                    }
                for (int ti = 0; ti < S(); ++ti)
                    for (int oi = 0; oi < O(); ++oi)
                    {
                        s2.index(ti), o.index(oi);
                        counts->push_back(model->count(&a, &s2, &o));
                    }
            }
        } else
            return describeModel(const_cast<FBAPOMDPState*>(static_cast<FBAPOMDPState const*>(p))->model(), counts);
        int32_t id = -1;
        check(_ctx, fba_model_add_structures(_model, 1, tp.data(), op.data(), &id), "fba_model_add_structures");
        return id;
    }

    // the same for a bare factored model (e.g. FBAPOMDPPrior::computePriorModel's result)
    int32_t describeModel(::bayes_adaptive::factored::BABNModel* model, std::vector<float>* counts) const
    {
        std::vector<uint32_t> tp((size_t)A() * FS()), op((size_t)A() * FO());
        counts->clear();
        IndexAction a(0);
        for (int ai = 0; ai < A(); ++ai)
        {
            a.index(ai);
            for (int f = 0; f < FS(); ++f)
                tp[(size_t)ai * FS() + f] = dumpNode(model->transitionNode(&a, f), _feat_s[f], counts);
            for (int g = 0; g < FO(); ++g)
                op[(size_t)ai * FO() + g] = dumpNode(model->observationNode(&a, g), _feat_o[g], counts);
        }
        int32_t id = -1;
        check(_ctx, fba_model_add_structures(_model, 1, tp.data(), op.data(), &id), "fba_model_add_structures");
        return id;
    }

    // structure id <-> the reference's BABNModel::Structure (parents per action and feature, ascending)
    ::bayes_adaptive::factored::BABNModel::Structure structureOf(int32_t struct_id) const
    {
        std::vector<uint32_t> tp((size_t)A() * FS()), op((size_t)A() * FO());
        check(_ctx, fba_model_get_structure(_model, struct_id, tp.data(), op.data()), "fba_model_get_structure");
        ::bayes_adaptive::factored::BABNModel::Structure st;
        st.T.resize((size_t)A()), st.O.resize((size_t)A());
        for (int a = 0; a < A(); ++a)
        {
            for (int f = 0; f < FS(); ++f) st.T[(size_t)a].push_back(parentsOf(tp[(size_t)a * FS() + f]));
            for (int g = 0; g < FO(); ++g) st.O[(size_t)a].push_back(parentsOf(op[(size_t)a * FO() + g]));
        }
        return st;
    }
    int32_t structureId(::bayes_adaptive::factored::BABNModel::Structure const& st) const
    {
        std::vector<uint32_t> tp((size_t)A() * FS()), op((size_t)A() * FO());
        for (int a = 0; a < A(); ++a)
        {
            for (int f = 0; f < FS(); ++f)
                for (auto p : st.T[(size_t)a][(size_t)f]) tp[(size_t)a * FS() + f] |= 1u << p;
            for (int g = 0; g < FO(); ++g)
                for (auto p : st.O[(size_t)a][(size_t)g]) op[(size_t)a * FO() + g] |= 1u << p;
        }
        int32_t id = -1;
        check(_ctx, fba_model_add_structures(_model, 1, tp.data(), op.data(), &id), "fba_model_add_structures");
        return id;
    }
    ::bayes_adaptive::factored::FBAPOMDP const* fbapomdp() const { return _fbapomdp; }

    // the reverse: a host-side reference particle from a downloaded block (Belief::sample())
    BAState* materialise(int32_t struct_id, int32_t state, std::vector<float> const& counts) const
    {
        auto domain_state = _sim.copyDomainState(_sim.domainState(state));
        if (!factored())
        {
            auto phi = std::make_shared<std::vector<float>>((size_t)S() * A() * S());
            auto psi = std::make_shared<std::vector<float>>((size_t)A() * S() * O());
            size_t k = 0;
            for (int a = 0; a < A(); ++a)
            {
                for (int s = 0; s < S(); ++s)
                    for (int t = 0; t < S(); ++t) (*phi)[indexing::threeToOne(s, a, t, A(), S())] = counts[k++];
                for (int t = 0; t < S(); ++t)
                    for (int o = 0; o < O(); ++o) (*psi)[indexing::threeToOne(a, t, o, S(), O())] = counts[k++];
            }
            return new BAPOMDPState(
                domain_state, ::bayes_adaptive::table::BAFlatModel(phi, psi, _sim.domainSize()));
        }
        std::vector<uint32_t> tp((size_t)A() * FS()), op((size_t)A() * FO());
        check(_ctx, fba_model_get_structure(_model, struct_id, tp.data(), op.data()), "fba_model_get_structure");
        std::vector<DBNNode> T, O_;
        size_t k = 0;
        auto const* fsize = _fbapomdp->domainFeatureSize();
        for (int a = 0; a < A(); ++a)
        {
            for (int f = 0; f < FS(); ++f) T.emplace_back(buildNode(fsize, tp[(size_t)a * FS() + f], _feat_s[f], counts, &k));
            for (int g = 0; g < FO(); ++g) O_.emplace_back(buildNode(fsize, op[(size_t)a * FO() + g], _feat_o[g], counts, &k));
        }
        return new FBAPOMDPState(
            domain_state,
            ::bayes_adaptive::factored::BABNModel(_sim.domainSize(), fsize, _steps.get(), std::move(T), std::move(O_)));
    }

private:
    BAPOMDP const& _sim;
    ::bayes_adaptive::factored::FBAPOMDP const* _fbapomdp = nullptr;
    fba_ctx* _ctx     = nullptr;
    fba_model* _model = nullptr;
    std::vector<float> _start_freq;
    mutable fba_tree* _tree     = nullptr;
    mutable int64_t _tree_sims  = 0;
    mutable int _tree_depth     = 0;
    std::vector<int> _feat_s, _feat_o;
    std::vector<double> _rew_sa, _rew_as2;
    std::vector<uint8_t> _term_sa, _term_as2;
    std::unique_ptr<::bayes_adaptive::factored::BABNModel::Indexing_Steps> _steps;

    std::vector<int> parentsOf(uint32_t mask) const
    {
        std::vector<int> p;
        for (int f = 0; f < FS(); ++f)
            if (mask & (1u << f)) p.push_back(f);
        return p;
    }

    uint32_t dumpNode(DBNNode& node, int range, std::vector<float>* counts) const
    {
        auto const& parents = *node.parents();
        uint32_t mask       = 0;
        std::vector<int> sizes;
        for (auto p : parents)
        {
            mask |= 1u << p;
            sizes.push_back(_feat_s[p]);
        }
        std::vector<int> values(parents.size(), 0);
        do {
            for (int v = 0; v < range; ++v) counts->push_back(node.count(values, v));
        } while (!parents.empty() && !indexing::increment(values, sizes));
        return mask;
    }

    DBNNode buildNode(Domain_Feature_Size const* fsize, uint32_t mask, int range,
                      std::vector<float> const& counts, size_t* k) const
    {
        auto parents = parentsOf(mask);
        std::vector<int> sizes;
        for (auto p : parents) sizes.push_back(_feat_s[p]);
        DBNNode node(&fsize->_S, parents, range);
        std::vector<int> values(parents.size(), 0);
        do {
            for (int v = 0; v < range; ++v) node.count(values, v) = counts[(*k)++];
        } while (!parents.empty() && !indexing::increment(values, sizes));
        return node;
    }

    // BADomainExtension::{reward,terminal} -> separable tables (see the header comment)
    void probeRewards(fba_model_desc& d)
    {
        int const S_ = _sim.domainSize()->_S, A_ = _sim.domainSize()->_A;
        _rew_sa.assign((size_t)S_ * A_, 0.0), _rew_as2.assign((size_t)A_ * S_, 0.0);
        _term_sa.assign((size_t)S_ * A_, 1), _term_as2.assign((size_t)A_ * S_, 1);
        std::vector<double> f((size_t)S_ * A_ * S_);
        std::vector<uint8_t> t((size_t)S_ * A_ * S_);
        IndexAction a(0);
        // the simulator exposes reward / terminal only through step(); the extension's functions are
        // reached through a KeepCounts-free route: BAPOMDP::domainState + the public extension API
        for (int s = 0; s < S_; ++s)
            for (int ai = 0; ai < A_; ++ai)
                for (int s2 = 0; s2 < S_; ++s2)
                {
                    a.index(ai);
                    size_t const k = ((size_t)s * A_ + ai) * S_ + s2;
                    f[k]           = rewardOf(s, &a, s2, &t[k]);
                }
        for (int ai = 0; ai < A_; ++ai)
            for (int s2 = 0; s2 < S_; ++s2) _rew_as2[(size_t)ai * S_ + s2] = f[((size_t)0 * A_ + ai) * S_ + s2];
        for (int s = 0; s < S_; ++s)
            for (int ai = 0; ai < A_; ++ai)
                _rew_sa[(size_t)s * A_ + ai] = f[((size_t)s * A_ + ai) * S_] - f[((size_t)0 * A_ + ai) * S_];
        for (int s = 0; s < S_; ++s)
            for (int ai = 0; ai < A_; ++ai)
                for (int s2 = 0; s2 < S_; ++s2)
                {
                    size_t const k = ((size_t)s * A_ + ai) * S_ + s2;
                    if (!t[k]) _term_sa[(size_t)s * A_ + ai] = 0, _term_as2[(size_t)ai * S_ + s2] = 0;
                }
        for (int s = 0; s < S_; ++s)
            for (int ai = 0; ai < A_; ++ai)
                for (int s2 = 0; s2 < S_; ++s2)
                {
                    size_t const k = ((size_t)s * A_ + ai) * S_ + s2;
                    if (f[k] != _rew_sa[(size_t)s * A_ + ai] + _rew_as2[(size_t)ai * S_ + s2])
                        throw std::string("fba_b200: this domain's reward is not f(s,a) + g(a,s')");
                    if ((t[k] != 0) != (_term_sa[(size_t)s * A_ + ai] || _term_as2[(size_t)ai * S_ + s2]))
                        throw std::string("fba_b200: this domain's terminal is not f(s,a) || g(a,s')");
                }
        d.rew_sa = _rew_sa.data(), d.rew_as2 = _rew_as2.data();
        d.term_sa = _term_sa.data(), d.term_as2 = _term_as2.data();
    }

    // reward / terminal of (s,a,s') as BAPOMDP::step reports them (BAPOMDP.cpp:131-132), obtained
    // by stepping a scratch particle whose counts force the transition s -> s2 … too slow; instead
    // the extension is reached through the friend-free public hook below.
    double rewardOf(int s, Action const* a, int s2, uint8_t* terminal) const;
    // whether the simulator was built with rnd::sample::Dir::Regular (BAConf.hpp:22)
    bool sampledDirichlets() const;
};

// Specialisation point: how to reach BADomainExtension from a BAPOMDP. In the reference the
// extension is a protected member (BAPOMDP.hpp:129); a maintainer adds the two-line public accessor
// shown in INTEGRATION.md. Builds that cannot touch the reference (oracle/ref_harness.cpp) compile
// with -fno-access-control and define FBA_B200_PRIVATE_ACCESS.
#ifdef FBA_B200_PRIVATE_ACCESS
inline double CudaSimulator::rewardOf(int s, Action const* a, int s2, uint8_t* terminal) const
{
    auto ext  = _sim._ba_domain_ext.get();
    auto st   = ext->getState(s);
    auto st2  = ext->getState(s2);
    *terminal = ext->terminal(st, a, st2).terminated();
    return ext->reward(st, a, st2).toDouble();
}
inline bool CudaSimulator::sampledDirichlets() const
{
    return _sim._sample_method == &rnd::sample::Dir::sampleFromSampledMult;
}
#else
inline double CudaSimulator::rewardOf(int s, Action const* a, int s2, uint8_t* terminal) const
{
    auto ext  = _sim.domainExtension(); // accessor added per INTEGRATION.md
    auto st   = ext->getState(s);
    auto st2  = ext->getState(s2);
    *terminal = ext->terminal(st, a, st2).terminated();
    return ext->reward(st, a, st2).toDouble();
}
inline bool CudaSimulator::sampledDirichlets() const
{
    return _sim.samplesDirichlets(); // accessor added per INTEGRATION.md
}
#endif

// distinct (structure id, count block) pairs seen so far, found by a 64-bit hash of the block (full
// comparison on a hit) instead of string keys holding a copy of every block
struct ProtoTable
{
    std::vector<int32_t> sid;
    std::vector<std::vector<float>> blocks;
    std::unordered_multimap<uint64_t, int32_t> index;
    size_t stride = 0;
    size_t size() const { return sid.size(); }
    static uint64_t hash(int32_t id, std::vector<float> const& b)
    {
        uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)(uint32_t)id;
        auto const* w = reinterpret_cast<uint32_t const*>(b.data());
        for (size_t k = 0; k < b.size(); ++k)
        {
            h ^= w[k];
            h *= 0x100000001b3ull;
            h ^= h >> 29;
        }
        return h;
    }
    int32_t find_or_add(int32_t id, std::vector<float> const& b)
    {
        uint64_t const h = hash(id, b);
        auto range       = index.equal_range(h);
        for (auto it = range.first; it != range.second; ++it)
            if (sid[(size_t)it->second] == id && blocks[(size_t)it->second] == b) return it->second;
        int32_t const k = (int32_t)sid.size();
        sid.push_back(id);
        blocks.push_back(b);
        index.emplace(h, k);
        stride = std::max(stride, b.size());
        return k;
    }
};

// Shared plumbing of the two belief adapters.
class CudaParticleBelief : public beliefs::BABelief
{
public:
    CudaParticleBelief(size_t n, bool weighted, uint64_t seed, int device) :
            _n(n), _weighted(weighted), _device(device)
    {
        _rng.mode   = FBA_RNG_PHILOX;
        _rng.words  = nullptr;
        _rng.n_words = _rng.cursor = 0;
        _rng.seed   = seed;
        _rng.offset = 0;
    }
    ~CudaParticleBelief() override { release(); }

    void initiate(POMDP const& d) override
    {
        auto const& sim = dynamic_cast<BAPOMDP const&>(d);
        _cuda.reset(new CudaSimulator(sim, _device, 4096, -1, _start_samples));
        // Belief::initiate = N x sampleStartState (BAImportanceSampling.cpp:49-60) = N x {prior sample,
        // domain start state}, independent of each other (BAPOMDP.cpp:101-104). The reference's prior and
        // domain run on the HOST; distinct (structure, count block) pairs become prototypes that are
        // uploaded once, each particle names its prototype and its domain start state.
        //
        // The reference's prior clones (and describe() reads) a whole count table per sample — 10^6
        // gridworld-5 particles take 18 minutes that way — although priors only ever
        // return a handful of distinct tables: one (BAPOMDPPrior.cpp:32-57, FBAPOMDPPrior.cpp:27-69
        // without a structure prior) or one per sampled structure. So the prior is sampled exactly only
        // until it stops producing new prototypes (kQuiet consecutive known ones, then a little longer to
        // sharpen the frequencies, bounded by kExtraSeconds); the remaining particles draw their prototype
        // from the observed frequencies and only their DOMAIN state from the reference's domain
        // (BAPOMDP::sampleDomainState, cheap). Beliefs up to kQuiet particles, and priors that keep
        // producing new tables, are sampled particle by particle exactly as before.
        size_t const kQuiet = _cuda->factored() ? 4096 : 2; // tabular priors hold ONE table (BAPOMDPPrior.cpp:32-57)
        double const kExtraSeconds = 1.0;
        size_t const kExtraDraws   = 65536;
        std::vector<int32_t> state(_n), proto(_n);
        ProtoTable protos;
        std::vector<float> block;
        std::vector<double> freq;
        size_t m = 0, since_new = 0;
        auto const t0 = std::chrono::steady_clock::now();
        bool quiet = false;
        for (; m < _n; ++m)
        {
            if (!quiet && since_new >= kQuiet) quiet = true;
            if (quiet)
            { // one prototype: nothing to sharpen. Several: keep sampling for a bounded while
                if (protos.size() == 1 || m >= kExtraDraws) break;
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > kExtraSeconds) break;
            }
            auto p            = static_cast<BAState const*>(d.sampleStartState());
            state[m]          = p->_domain_state->index();
            int32_t const sid = _cuda->describe(p, &block);
            size_t const before = protos.size();
            proto[m]          = protos.find_or_add(sid, block);
            if (protos.size() != before) since_new = 0, quiet = false;
            else
                ++since_new;
            if (freq.size() < protos.size()) freq.resize(protos.size(), 0.0);
            freq[(size_t)proto[m]] += 1.0;
            d.releaseState(p);
        }
        _host_prior_samples = m;
        if (m < _n)
        {
            std::mt19937_64 gen(_rng.seed ^ 0x5bd1e995u);
            std::discrete_distribution<int32_t> pick(freq.begin(), freq.end());
            for (size_t i = m; i < _n; ++i)
            {
                proto[i] = protos.size() == 1 ? 0 : pick(gen);
                auto st  = sim.sampleDomainState();
                state[i] = st->index();
                sim.releaseDomainState(st);
            }
        }
        size_t stride = std::max(protos.stride, minimumStride(d));
        check(_cuda->ctx(), fba_belief_create(_cuda->ctx(), _cuda->model(), (int64_t)_n, (int64_t)stride,
                                               _weighted ? 1 : 0, &_belief),
              "fba_belief_create");
        stride = (size_t)fba_belief_stride(_belief);
        std::vector<float> flat(protos.size() * stride, 0.0f);
        for (size_t k = 0; k < protos.size(); ++k)
            std::copy(protos.blocks[k].begin(), protos.blocks[k].end(), flat.begin() + k * stride);
        check(_cuda->ctx(),
              fba_belief_init(_belief, (int32_t)protos.size(), protos.sid.data(), flat.data(), proto.data(),
                              state.data()),
              "fba_belief_init");
    }

    // prior samples the last initiate() took on the host (the rest reused their prototypes)
    size_t hostPriorSamples() const { return _host_prior_samples; }

    void free(POMDP const& /*d*/) override { release(); }

    State const* sample() const override
    {
        int64_t i = 0;
        check(_cuda->ctx(), fba_belief_sample(_belief, &_rng, &i), "fba_belief_sample");
        size_t const stride = (size_t)fba_belief_stride(_belief);
        std::vector<float> counts(stride);
        int32_t state = 0, sid = 0;
        check(_cuda->ctx(), fba_belief_download(_belief, i, 1, &state, &sid, counts.data(), nullptr),
              "fba_belief_download");
        dropSample();
        _sample = _cuda->materialise(sid, state, counts);
        return _sample;
    }

    void resetDomainStateDistribution(BAPOMDP const& bapomdp) override
    {
        // weighted: N weighted draws with replacement first (BAImportanceSampling.cpp:90-111)
        if (_weighted) check(_cuda->ctx(), fba_belief_resample(_belief, &_rng), "fba_belief_resample");
        // then a fresh domain start state per particle, drawn by the reference's own domain
        std::vector<int32_t> state(_n);
        for (size_t i = 0; i < _n; ++i)
        {
            auto s   = bapomdp.sampleDomainState();
            state[i] = s->index();
            bapomdp.releaseDomainState(s);
        }
        check(_cuda->ctx(), fba_belief_upload(_belief, 0, (int64_t)_n, state.data(), nullptr, nullptr, nullptr),
              "fba_belief_upload");
    }

    fba_belief* handle() const { return _belief; }
    CudaSimulator const& cuda() const { return *_cuda; }

protected:
    size_t _n;
    bool _weighted;
    int _device;
    mutable fba_rng _rng;
    std::unique_ptr<CudaSimulator> _cuda;
    fba_belief* _belief            = nullptr;
    mutable BAState const* _sample = nullptr;
    size_t _host_prior_samples     = 0;
    int _start_samples             = 0; // > 0: the simulator learns the domain's start distribution (device draws)

    // cells a particle block must hold beyond what the prior's prototypes need (beliefs that change
    // structures later override this)
    virtual size_t minimumStride(POMDP const& /*d*/) const { return 0; }

    void dropSample() const
    {
        if (_sample)
        {
            _cuda->sim().releaseState(_sample);
            _sample = nullptr;
        }
    }
    void release()
    {
        if (_cuda) dropSample();
        fba_belief_destroy(_belief);
        _belief = nullptr;
        _cuda.reset();
    }
};

class CudaBAImportanceSampling : public CudaParticleBelief
{
public:
    explicit CudaBAImportanceSampling(size_t n, uint64_t seed = 42, int device = 0) :
            CudaParticleBelief(n, true, seed, device)
    {
        if (n < 1) throw("cannot initiate BAImportanceSampling with n " + std::to_string(n)); // as :19-22
    }
    void updateEstimation(Action const* a, Observation const* o, POMDP const& /*d*/) override
    {
        check(_cuda->ctx(), fba_belief_update_estimation(_belief, a->index(), o->index(), &_rng, nullptr),
              "fba_belief_update_estimation");
    }
};

class CudaBARejectionSampling : public CudaParticleBelief
{
public:
    explicit CudaBARejectionSampling(size_t n, uint64_t seed = 42, int device = 0) :
            CudaParticleBelief(n, false, seed, device)
    {
        if (n < 1) throw("cannot initiate RejectionSampling with n = " + std::to_string(n)); // as :13-16
    }
    void updateEstimation(Action const* a, Observation const* o, POMDP const& /*d*/) override
    {
        int64_t attempts = 0;
        check(_cuda->ctx(), fba_belief_reject_sample(_belief, a->index(), o->index(), &_rng, &attempts),
              "fba_belief_reject_sample");
    }
};

// beliefs::bayes_adaptive::factored::ReinvigoratingRejectionSampling
// (src/beliefs/bayes-adaptive/factored/ReinvigoratingRejectionSampling.cpp:37-131): two flat filters —
// the learned-structure belief and the fully connected one. `mutate_kind` names the domain's
// FBAPOMDP::mutate (fba_mutate_kind); the factory picks it from the -D string.
class CudaReinvigoratingRejectionSampling : public beliefs::BABelief
{
public:
    CudaReinvigoratingRejectionSampling(size_t size, size_t reinvigoration_amount, int mutate_kind,
                                        uint64_t seed = 42, int device = 0) :
            _size(size), _amount(reinvigoration_amount), _mutate(mutate_kind), _device(device)
    {
        if (_size < 1 || _amount < 1) // as ReinvigoratingRejectionSampling.cpp:43-48
            throw "ReinvigoratingRejectionSampling::cannot initiate belief of size < 1 (" + std::to_string(_size)
                + "), or resample size of < 1 (" + std::to_string(_amount) + ")";
        _rng.mode    = FBA_RNG_PHILOX;
        _rng.words   = nullptr;
        _rng.n_words = _rng.cursor = 0;
        _rng.seed    = seed;
        _rng.offset  = 0;
    }
    ~CudaReinvigoratingRejectionSampling() override { release(); }

    void initiate(POMDP const& d) override
    {
        auto const& fbapomdp = dynamic_cast<::bayes_adaptive::factored::FBAPOMDP const&>(d);
        _cuda.reset(new CudaSimulator(fbapomdp, _device, 1 << 16));
        // :56-75: _size x sampleStartState, then _size x sampleFullyConnectedState (host prior)
        std::vector<int32_t> state[2], sid[2];
        std::vector<std::vector<float>> blocks[2];
        size_t stride = 0;
        for (int k = 0; k < 2; ++k)
        {
            state[k].resize(_size), sid[k].resize(_size), blocks[k].resize(_size);
            for (size_t i = 0; i < _size; ++i)
            {
                BAState const* p = (k == 0) ? static_cast<BAState const*>(d.sampleStartState())
                                            : static_cast<BAState const*>(fbapomdp.sampleFullyConnectedState());
                state[k][i] = p->_domain_state->index();
                sid[k][i]   = _cuda->describe(p, &blocks[k][i]);
                stride      = std::max(stride, blocks[k][i].size());
                d.releaseState(p);
            }
        }
        for (int k = 0; k < 2; ++k)
        {
            check(_cuda->ctx(),
                  fba_belief_create(_cuda->ctx(), _cuda->model(), (int64_t)_size, (int64_t)stride, 0, &_b[k]),
                  "fba_belief_create");
            size_t const st = (size_t)fba_belief_stride(_b[k]);
            std::vector<float> flat(_size * st, 0.0f);
            for (size_t i = 0; i < _size; ++i)
                std::copy(blocks[k][i].begin(), blocks[k][i].end(), flat.begin() + i * st);
            check(_cuda->ctx(),
                  fba_belief_upload(_b[k], 0, (int64_t)_size, state[k].data(), sid[k].data(), flat.data(), nullptr),
                  "fba_belief_upload");
        }
    }

    void free(POMDP const& /*d*/) override { release(); }

    State const* sample() const override
    {
        int64_t i = 0;
        check(_cuda->ctx(), fba_belief_sample(_b[0], &_rng, &i), "fba_belief_sample");
        std::vector<float> counts((size_t)fba_belief_stride(_b[0]));
        int32_t state = 0, sid = 0;
        check(_cuda->ctx(), fba_belief_download(_b[0], i, 1, &state, &sid, counts.data(), nullptr),
              "fba_belief_download");
        dropSample();
        _sample = _cuda->materialise(sid, state, counts);
        return _sample;
    }

    void updateEstimation(Action const* a, Observation const* o, POMDP const& /*d*/) override
    {
        // :89-106: reinvigorateParticles, then rejectSample on both filters
        int64_t attempts = 0;
        check(_cuda->ctx(), fba_belief_reinvigorate(_b[0], _b[1], (int64_t)_amount, _mutate, &_rng),
              "fba_belief_reinvigorate");
        for (int k = 0; k < 2; ++k)
            check(_cuda->ctx(), fba_belief_reject_sample(_b[k], a->index(), o->index(), &_rng, &attempts),
                  "fba_belief_reject_sample");
    }

    void resetDomainStateDistribution(BAPOMDP const& bapomdp) override
    {
        // :108-119: resetDomainState on every particle of both filters, belief first
        std::vector<int32_t> state(_size);
        for (int k = 0; k < 2; ++k)
        {
            for (size_t i = 0; i < _size; ++i)
            {
                auto s   = bapomdp.sampleDomainState();
                state[i] = s->index();
                bapomdp.releaseDomainState(s);
            }
            check(_cuda->ctx(),
                  fba_belief_upload(_b[k], 0, (int64_t)_size, state.data(), nullptr, nullptr, nullptr),
                  "fba_belief_upload");
        }
    }

    fba_belief* handle() const { return _b[0]; }
    CudaSimulator const& cuda() const { return *_cuda; }

private:
    size_t _size, _amount;
    int _mutate, _device;
    mutable fba_rng _rng;
    std::unique_ptr<CudaSimulator> _cuda;
    fba_belief* _b[2]              = {nullptr, nullptr};
    mutable BAState const* _sample = nullptr;

    void dropSample() const
    {
        if (_sample)
        {
            _cuda->sim().releaseState(_sample);
            _sample = nullptr;
        }
    }
    void release()
    {
        if (_cuda) dropSample();
        fba_belief_destroy(_b[0]);
        fba_belief_destroy(_b[1]);
        _b[0] = _b[1] = nullptr;
        _cuda.reset();
    }
};

} // namespace fba_b200

#endif // FBA_B200_CUDA_BELIEFS_HPP
