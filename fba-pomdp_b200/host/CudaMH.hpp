// Drop-in for beliefs::bayes_adaptive::factored::MHNIPS2018
// (src/beliefs/bayes-adaptive/factored/MHNIPS2018.{hpp,cpp}), the first of the reference's
// Metropolis-Hastings structure beliefs, behind the same BABelief interface.
//
// What runs where:
//   * updateEstimation (MHNIPS2018.cpp:166-186) = importance-sampling update + resample on the GPU
//     (fba_belief_update_estimation, the kernels of SURVEY.md §8 a1/a2), the step likelihood read back to
//     accumulate the log likelihood on the host, the (action, observation) history kept on the host.
//   * MH (MHNIPS2018.cpp:188-255), triggered when the log likelihood drops below the threshold: the
//     reference proposes ONE structure at a time and replays the whole history on it before the next
//     proposal; proposals are independent of each other, so here a BATCH of proposals is drawn at once —
//     source particles on the device (fba_belief_sample_batch), the 50/50 keep-or-mutate choice, the
//     domain's own `mutate` and the prior model of each proposed structure by the reference's own code on
//     the host (FBAPOMDP::mutate, FBAPOMDPPrior::computePriorModel: small, domain-specific) — and then
//     every proposal replays the history in its own GPU thread (fba_belief_replay_history =
//     computePosterior, MHNIPS2018.cpp:41-109) and is scored against its prior with
//     BABNModel::LogBDScore on the GPU (fba_belief_log_bd_score). Accept / reject (MHNIPS2018.cpp:240) is
//     one comparison per proposal on the host; accepted proposals, in proposal order, fill the new belief
//     (fba_belief_assign_from) until it holds `size` particles — the distribution of the new belief is the
//     reference's, because its proposals are i.i.d. too.
#ifndef FBA_B200_CUDA_MH_HPP
#define FBA_B200_CUDA_MH_HPP

#include <cmath>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "CudaBeliefs.hpp"

#include "bayes-adaptive/priors/FBAPOMDPPrior.hpp"

namespace fba_b200 {

// The prior model of a structure (FBAPOMDPPrior::computePriorModel, by the reference's own code) in this repo's block
// layout, cached per structure id for the duration of one MH / Gibbs run, plus the id <-> Structure maps both
// samplers need. `flat` / `sid` are what fba_belief_init takes as prototypes.
class PriorModelCache
{
public:
    using Structure = ::bayes_adaptive::factored::BABNModel::Structure;

    PriorModelCache(CudaSimulator const& cuda, ::bayes_adaptive::factored::FBAPOMDP const& fbapomdp, int64_t stride,
                    char const* who) :
            _cuda(cuda), _fbapomdp(fbapomdp), _stride(stride), _who(who)
    {
    }

    // prototype index of structure `id` (computing and caching its prior model on first use)
    int32_t of(int32_t id, Structure const& st)
    {
        auto it = _index.find(id);
        if (it != _index.end()) return it->second;
        auto model        = _fbapomdp.prior()->computePriorModel(st);
        int32_t const got = _cuda.describeModel(&model, &_block);
        if (got != id) throw std::string(_who) + ": the prior model of a structure has another structure";
        if ((int64_t)_block.size() > _stride) throw std::string(_who) + ": structure larger than the particle blocks";
        int32_t const k = (int32_t)sid.size();
        sid.push_back(id);
        flat.resize((size_t)(k + 1) * _stride, 0.0f);
        std::copy(_block.begin(), _block.end(), flat.begin() + (size_t)k * _stride);
        _index.emplace(id, k);
        return k;
    }
    int32_t of(int32_t id) { return of(id, structure(id)); }

    Structure const& structure(int32_t id)
    {
        auto it = _structures.find(id);
        if (it == _structures.end()) it = _structures.emplace(id, _cuda.structureOf(id)).first;
        return it->second;
    }

    // every particle of b becomes the prior model proto[i] names (zeros: a vector of at least b's size)
    void setPriors(fba_belief* b, std::vector<int32_t> const& proto, std::vector<int32_t> const& zeros) const
    {
        check(_cuda.ctx(), fba_belief_init(b, (int32_t)sid.size(), sid.data(), flat.data(), proto.data(), zeros.data()),
              "fba_belief_init");
    }

    std::vector<int32_t> sid; // prototype k has structure sid[k] ...
    std::vector<float> flat;  // ... and counts flat[k * stride ..]

private:
    CudaSimulator const& _cuda;
    ::bayes_adaptive::factored::FBAPOMDP const& _fbapomdp;
    int64_t _stride;
    char const* _who;
    std::vector<float> _block;
    std::map<int32_t, int32_t> _index;
    std::map<int32_t, Structure> _structures;
};

class CudaMHNIPS2018 : public CudaParticleBelief
{
public:
    CudaMHNIPS2018(size_t size, double ll_threshold, uint64_t seed = 42, int device = 0) :
            CudaParticleBelief(size, true, seed, device), _ll_threshold(ll_threshold)
    {
        if (size < 1) throw("MHNIPS2018::cannot initiate MH with size " + std::to_string(size)); // as :114-117
        if (ll_threshold >= 0)                                                                    // as :119-124
            throw("MHNIPS2018::cannot initiate with threshold >= 0 (is:" + std::to_string(ll_threshold) + ")");
        _start_samples = 1 << 16; // computePosterior draws a domain start state per episode attempt, on the device
    }

    void initiate(POMDP const& d) override
    {
        CudaParticleBelief::initiate(d);
        _history.assign(1, {});
        _log_likelihood = 0.0;
        _mh_runs = _proposals = 0;
    }

    // MHNIPS2018.cpp:132-147: a fresh domain state for every particle (no resampling: the weights are
    // uniform after every update), and a new episode in the history
    void resetDomainStateDistribution(BAPOMDP const& bapomdp) override
    {
        std::vector<int32_t> state(_n);
        for (size_t i = 0; i < _n; ++i)
        {
            auto s   = bapomdp.sampleDomainState();
            state[i] = s->index();
            bapomdp.releaseDomainState(s);
        }
        check(_cuda->ctx(), fba_belief_upload(_belief, 0, (int64_t)_n, state.data(), nullptr, nullptr, nullptr),
              "fba_belief_upload");
        if (!_history.back().empty()) _history.emplace_back();
    }

    void updateEstimation(Action const* a, Observation const* o, POMDP const& d) override
    {
        double lik = 0.0;
        check(_cuda->ctx(), fba_belief_update_estimation(_belief, a->index(), o->index(), &_rng, &lik),
              "fba_belief_update_estimation");
        _log_likelihood += std::log(lik); // :169
        _history.back().emplace_back(a->index(), o->index());
        if (_log_likelihood < _ll_threshold) MH(d); // :175-178
    }

    double logLikelihood() const { return _log_likelihood; }
    size_t mhRuns() const { return _mh_runs; }
    size_t proposals() const { return _proposals; }

protected:
    // MH may propose any structure the domain's mutate can reach: blocks are sized for the fully
    // connected one (FBAPOMDP::sampleFullyConnectedState, as the reinvigoration belief does)
    size_t minimumStride(POMDP const& d) const override
    {
        auto const& fbapomdp = dynamic_cast<::bayes_adaptive::factored::FBAPOMDP const&>(d);
        auto p               = static_cast<BAState const*>(fbapomdp.sampleFullyConnectedState());
        std::vector<float> block;
        _cuda->describe(p, &block);
        d.releaseState(p);
        return block.size();
    }

private:
    double _ll_threshold;
    double _log_likelihood = 0.0;
    std::vector<std::vector<std::pair<int32_t, int32_t>>> _history; // [episode][step] = (action, observation)
    size_t _mh_runs = 0, _proposals = 0;

    struct Beliefs // proposal scratch, released on every exit path
    {
        std::vector<fba_belief*> all;
        ~Beliefs()
        {
            for (auto b : all) fba_belief_destroy(b);
        }
        fba_belief* make(fba_ctx* ctx, fba_model* m, int64_t n, int64_t stride, int weighted)
        {
            fba_belief* b = nullptr;
            check(ctx, fba_belief_create(ctx, m, n, stride, weighted, &b), "fba_belief_create");
            all.push_back(b);
            return b;
        }
        void drop(fba_belief* b)
        {
            for (auto& x : all)
                if (x == b) x = nullptr;
            fba_belief_destroy(b);
        }
    };

    void MH(POMDP const& d)
    {
        auto const& fbapomdp = dynamic_cast<::bayes_adaptive::factored::FBAPOMDP const&>(d);
        fba_ctx* ctx         = _cuda->ctx();
        int64_t const N      = (int64_t)_n;
        int64_t const stride = fba_belief_stride(_belief);
        Beliefs tmp;

        PriorModelCache priors(*_cuda, fbapomdp, stride, "CudaMHNIPS2018");
        auto priorBelief = [&](std::vector<int32_t> const& proto) { // n particles, particle i = prior prototype proto[i]
            fba_belief* b = tmp.make(ctx, _cuda->model(), (int64_t)proto.size(), stride, 0);
            priors.setPriors(b, proto, std::vector<int32_t>(proto.size(), 0));
            return b;
        };

        // old_score of every particle: LogBDScore against the prior model of ITS structure (:237)
        std::vector<int32_t> sid((size_t)N), proto((size_t)N);
        check(ctx, fba_belief_download(_belief, 0, N, nullptr, sid.data(), nullptr, nullptr), "fba_belief_download");
        for (int64_t i = 0; i < N; ++i) proto[(size_t)i] = priors.of(sid[(size_t)i]);
        std::vector<double> old_score((size_t)N);
        {
            fba_belief* pb = priorBelief(proto);
            check(ctx, fba_belief_log_bd_score(_belief, pb, old_score.data()), "fba_belief_log_bd_score");
            tmp.drop(pb);
        }

        // the history, flattened
        std::vector<int32_t> len, act, obs;
        for (auto const& ep : _history)
        {
            len.push_back((int32_t)ep.size());
            for (auto const& st : ep) act.push_back(st.first), obs.push_back(st.second);
        }

        fba_belief* fresh = tmp.make(ctx, _cuda->model(), N, stride, 1);
        {
            std::vector<int32_t> zeros((size_t)N, 0);
            check(ctx, fba_belief_init(fresh, 1, priors.sid.data(), priors.flat.data(), zeros.data(), zeros.data()),
                  "fba_belief_init"); // placeholder particles, uniform weights 1/N (:243); every slot is overwritten
        }
        int64_t accepted = 0;
        // proposal scratch, created once per MH run (a belief is ~20 device allocations) and re-initialised per batch
        int64_t const P   = std::max<int64_t>(N, 256);
        fba_belief* prop  = tmp.make(ctx, _cuda->model(), P, stride, 0);
        fba_belief* prior = tmp.make(ctx, _cuda->model(), P, stride, 0);
        std::vector<int32_t> zerosP((size_t)P, 0);
        auto setPriors = [&](fba_belief* b, std::vector<int32_t> const& proto) { priors.setPriors(b, proto, zerosP); };
        while (accepted < N)
        {
            std::vector<int64_t> src((size_t)P);
            check(ctx, fba_belief_sample_batch(_belief, &_rng, P, src.data()), "fba_belief_sample_batch"); // :200
            std::vector<int32_t> pproto((size_t)P);
            for (int64_t j = 0; j < P; ++j)
            {
                int32_t const id = sid[(size_t)src[(size_t)j]];
                if (rnd::boolean()) pproto[(size_t)j] = priors.of(id); // :206-208: same structure half the time
                else
                {
                    auto st            = fbapomdp.mutate(priors.structure(id));
                    int32_t const nid  = _cuda->structureId(st);
                    pproto[(size_t)j]  = priors.of(nid, st);
                }
            }
            setPriors(prop, pproto);
            setPriors(prior, pproto);
            check(ctx,
                  fba_belief_replay_history(prop, (int32_t)len.size(), len.data(), act.data(), obs.data(), &_rng,
                                            1000000),
                  "fba_belief_replay_history"); // :214-215
            std::vector<double> new_score((size_t)P);
            check(ctx, fba_belief_log_bd_score(prop, prior, new_score.data()), "fba_belief_log_bd_score"); // :238
            std::vector<int64_t> take;
            for (int64_t j = 0; j < P && accepted + (int64_t)take.size() < N; ++j)
                if (std::log(rnd::uniform_rand01()) < new_score[(size_t)j] - old_score[(size_t)src[(size_t)j]]) // :240
                    take.push_back(j);
            check(ctx, fba_belief_assign_from(fresh, accepted, prop, (int64_t)take.size(), take.data()),
                  "fba_belief_assign_from");
            accepted += (int64_t)take.size();
            _proposals += (size_t)P;
        }
        // the new belief replaces the old one (:252)
        for (auto& x : tmp.all)
            if (x == fresh) x = nullptr;
        dropSample();
        fba_belief_destroy(_belief);
        _belief         = fresh;
        _log_likelihood = 0.0; // :253
        ++_mh_runs;
    }
};

// Drop-in for beliefs::bayes_adaptive::factored::MHwithinGibbs
// (src/beliefs/bayes-adaptive/factored/MHwithinGibbs.{hpp,cpp}): importance sampling as above, and when the log
// likelihood drops below the threshold a Gibbs sampler over (state history, model) rebuilds the belief
// (reinvigorate, MHwithinGibbs.cpp:334-395):
//   * p(states | model): fba_belief_sample_state_history — backward messages + forward sampling over the
//     flattened model (MSG), or rejection sampling (RS) — on the GPU;
//   * p(model | states): Metropolis-Hastings over structures — the domain's mutate and the prior model of the
//     proposed structure by the reference's own code on the host, computePosteriorCounts
//     (fba_belief_add_history_counts) and BABNModel::LogBDScore (fba_belief_log_bd_score) on the GPU.
// The reference runs ONE chain until it has produced `size` models: strictly sequential, a handful of tiny
// operations per link. Here `chains` independent chains advance in LOCKSTEP, each started from its own draw of
// the old belief and each following the reference's transition exactly (including its detail that the new state
// history is sampled from the model BEFORE the move, :377-378): one batched call per operation serves every
// chain, and the accepted models of all chains, in (sweep, chain) order, fill the new belief. chains = 1 is the
// reference's algorithm link for link; more chains change the mixing (for the better: independent starts), not
// the transition kernel. Default: min(size, 256).
class CudaMHwithinGibbs : public CudaParticleBelief
{
public:
    enum SAMPLE_STATE_HISTORY_TYPE { RS, MSG }; // as MHwithinGibbs.hpp

    CudaMHwithinGibbs(size_t size, double ll_threshold, SAMPLE_STATE_HISTORY_TYPE type, uint64_t seed = 42,
                      int device = 0, size_t chains = 0) :
            CudaParticleBelief(size, true, seed, device),
            _ll_threshold(ll_threshold),
            _type(type),
            _n_chains(chains ? std::min(chains, size) : std::min<size_t>(size, 256))
    {
        if (size < 1) throw "MHwithinGibbs::cannot initiate MH with size 0"; // as :242-245
        if (ll_threshold >= 0)                                                // as :247-251
            throw("MHwithinGibbs::cannot initiate with threshold >= 0 (is:" + std::to_string(ll_threshold) + ")");
    }

    void initiate(POMDP const& d) override
    {
        CudaParticleBelief::initiate(d);
        _history.assign(1, {});
        _log_likelihood = 0.0;
        _chains = _proposals = _sweeps = 0;
    }

    // :257-272: a fresh domain state for every particle where it is, and a new episode in the history
    void resetDomainStateDistribution(BAPOMDP const& bapomdp) override
    {
        std::vector<int32_t> state(_n);
        for (size_t i = 0; i < _n; ++i)
        {
            auto s   = bapomdp.sampleDomainState();
            state[i] = s->index();
            bapomdp.releaseDomainState(s);
        }
        check(_cuda->ctx(), fba_belief_upload(_belief, 0, (int64_t)_n, state.data(), nullptr, nullptr, nullptr),
              "fba_belief_upload");
        if (!_history.back().empty()) _history.emplace_back();
    }

    void updateEstimation(Action const* a, Observation const* o, POMDP const& d) override
    { // :310-327
        double lik = 0.0;
        check(_cuda->ctx(), fba_belief_update_estimation(_belief, a->index(), o->index(), &_rng, &lik),
              "fba_belief_update_estimation");
        _log_likelihood += std::log(lik);
        _history.back().emplace_back(a->index(), o->index());
        if (_log_likelihood < _ll_threshold) reinvigorate(d);
    }

    double logLikelihood() const { return _log_likelihood; }
    size_t chains() const { return _chains; }       // reinvigorations run
    size_t proposals() const { return _proposals; } // structure proposals scored
    size_t sweeps() const { return _sweeps; }       // lockstep sweeps over the chains

protected:
    size_t minimumStride(POMDP const& d) const override
    {
        auto const& fbapomdp = dynamic_cast<::bayes_adaptive::factored::FBAPOMDP const&>(d);
        auto p               = static_cast<BAState const*>(fbapomdp.sampleFullyConnectedState());
        std::vector<float> block;
        _cuda->describe(p, &block);
        d.releaseState(p);
        return block.size();
    }

private:
    double _ll_threshold;
    SAMPLE_STATE_HISTORY_TYPE _type;
    size_t _n_chains;
    double _log_likelihood = 0.0;
    std::vector<std::vector<std::pair<int32_t, int32_t>>> _history;
    size_t _chains = 0, _proposals = 0, _sweeps = 0;

    struct Scratch // beliefs that die with the call, whichever way it ends
    {
        std::vector<fba_belief*> all;
        ~Scratch()
        {
            for (auto b : all) fba_belief_destroy(b);
        }
        fba_belief* make(fba_ctx* ctx, fba_model* m, int64_t n, int64_t stride, int weighted)
        {
            fba_belief* b = nullptr;
            check(ctx, fba_belief_create(ctx, m, n, stride, weighted, &b), "fba_belief_create");
            all.push_back(b);
            return b;
        }
        void keep(fba_belief* b)
        {
            for (auto& x : all)
                if (x == b) x = nullptr;
        }
    };

    void reinvigorate(POMDP const& d)
    {
        auto const& fbapomdp = dynamic_cast<::bayes_adaptive::factored::FBAPOMDP const&>(d);
        fba_ctx* ctx         = _cuda->ctx();
        int64_t const N      = (int64_t)_n;
        int64_t const C      = (int64_t)_n_chains;
        int64_t const stride = fba_belief_stride(_belief);
        Scratch tmp;

        // the history, flattened; the prior of s_0
        std::vector<int32_t> len, act, obs;
        for (auto const& ep : _history)
        {
            len.push_back((int32_t)ep.size());
            for (auto const& st : ep) act.push_back(st.first), obs.push_back(st.second);
        }
        int64_t const L = (int64_t)act.size() + (int64_t)len.size();
        std::vector<float> state_prior((size_t)_cuda->S());
        for (int s = 0; s < _cuda->S(); ++s) state_prior[(size_t)s] = fbapomdp.domainStatePrior()->prob((size_t)s);

        PriorModelCache priors(*_cuda, fbapomdp, stride, "CudaMHwithinGibbs");
        std::vector<int32_t> zeros((size_t)std::max(C, N), 0);
        // every particle of b becomes the prior model its chain names
        auto setPriors = [&](fba_belief* b, std::vector<int32_t> const& proto) { priors.setPriors(b, proto, zeros); };
        auto sampleHistories = [&](fba_belief* models, std::vector<int32_t>& seq) { // sampleStateHistory, :215-232
            seq.resize((size_t)(C * L));
            check(ctx,
                  fba_belief_sample_state_history(models, _type == MSG ? 0 : 1, (int32_t)len.size(), len.data(), act.data(),
                                                  obs.data(), state_prior.data(), &_rng, 1000000, seq.data()),
                  "fba_belief_sample_state_history");
        };
        auto addCounts = [&](fba_belief* b, std::vector<int32_t> const& seq) { // computePosteriorCounts, :397-436
            check(ctx, fba_belief_add_history_counts(b, (int32_t)len.size(), len.data(), act.data(), obs.data(), seq.data(), 0),
                  "fba_belief_add_history_counts");
        };

        fba_belief* cur  = tmp.make(ctx, _cuda->model(), C, stride, 0); // every chain's current model
        fba_belief* prop = tmp.make(ctx, _cuda->model(), C, stride, 0); // proposals / priors, by turns
        fba_belief* pri  = tmp.make(ctx, _cuda->model(), C, stride, 0);
        std::vector<int64_t> all((size_t)C);
        for (int64_t c = 0; c < C; ++c) all[(size_t)c] = c;

        // :344-352 per chain: a particle of the old belief, a state history from its model, the posterior counts of
        // its structure's prior on that history, and their score
        std::vector<int64_t> start((size_t)C);
        check(ctx, fba_belief_sample_batch(_belief, &_rng, C, start.data()), "fba_belief_sample_batch");
        check(ctx, fba_belief_assign_from(cur, 0, _belief, C, start.data()), "fba_belief_assign_from");
        std::vector<int32_t> seq, new_seq, cur_sid((size_t)C), cur_proto((size_t)C);
        sampleHistories(cur, seq);
        check(ctx, fba_belief_download(cur, 0, C, nullptr, cur_sid.data(), nullptr, nullptr), "fba_belief_download");
        for (int64_t c = 0; c < C; ++c) cur_proto[(size_t)c] = priors.of(cur_sid[(size_t)c]);
        setPriors(pri, cur_proto);
        addCounts(pri, seq);
        check(ctx, fba_belief_replace_from(cur, all.data(), pri, all.data(), C), "fba_belief_replace_from");
        setPriors(prop, cur_proto);
        std::vector<double> score((size_t)C), new_score((size_t)C);
        check(ctx, fba_belief_log_bd_score(cur, prop, score.data()), "fba_belief_log_bd_score");

        fba_belief* fresh = tmp.make(ctx, _cuda->model(), N, stride, 1);
        check(ctx, fba_belief_init(fresh, 1, priors.sid.data(), priors.flat.data(), zeros.data(), zeros.data()),
              "fba_belief_init"); // placeholders with weight 1 / N (:372-374); every slot is overwritten
        int64_t accepted = 0;
        std::vector<int32_t> pproto((size_t)C), pid((size_t)C);
        while (accepted < N) // the Gibbs loop, :356-391, one link of every chain per sweep
        {
            ++_sweeps;
            // :360-365: every chain's proposal mutate(model.structure()), its posterior counts and score
            for (int64_t c = 0; c < C; ++c)
            {
                auto st           = fbapomdp.mutate(priors.structure(cur_sid[(size_t)c]));
                pid[(size_t)c]    = _cuda->structureId(st);
                pproto[(size_t)c] = priors.of(pid[(size_t)c], st);
            }
            setPriors(prop, pproto);
            setPriors(pri, pproto);
            addCounts(prop, seq);
            check(ctx, fba_belief_log_bd_score(prop, pri, new_score.data()), "fba_belief_log_bd_score");
            _proposals += (size_t)C;
            std::vector<int64_t> took;
            for (int64_t c = 0; c < C; ++c)
                if (std::log(rnd::uniform_rand01()) < new_score[(size_t)c] - score[(size_t)c]) took.push_back(c); // :367
            if (took.empty()) continue;
            // :370-374: the accepted models join the belief with the last state of their chain's CURRENT history
            int64_t const n_new = std::min<int64_t>((int64_t)took.size(), N - accepted);
            std::vector<int32_t> last((size_t)n_new);
            for (int64_t k = 0; k < n_new; ++k) last[(size_t)k] = seq[(size_t)(took[(size_t)k] * L + L - 1)];
            check(ctx, fba_belief_assign_from(fresh, accepted, prop, n_new, took.data()), "fba_belief_assign_from");
            check(ctx, fba_belief_upload(fresh, accepted, n_new, last.data(), nullptr, nullptr, nullptr), "fba_belief_upload");
            accepted += n_new;
            if (accepted >= N) break;
            // :377-383: a new state history from the model BEFORE the move, then the model of the accepted structure
            // on that history and its score (all chains are sampled; only the moved ones take the result)
            sampleHistories(cur, new_seq);
            for (auto c : took)
            {
                std::copy(new_seq.begin() + c * L, new_seq.begin() + (c + 1) * L, seq.begin() + c * L);
                cur_sid[(size_t)c]   = pid[(size_t)c];
                cur_proto[(size_t)c] = pproto[(size_t)c];
            }
            addCounts(pri, seq);                 // pri holds every chain's proposed prior: now prior + counts(history)
            check(ctx, fba_belief_replace_from(cur, took.data(), pri, took.data(), (int64_t)took.size()),
                  "fba_belief_replace_from");
            setPriors(prop, cur_proto);
            check(ctx, fba_belief_log_bd_score(cur, prop, new_score.data()), "fba_belief_log_bd_score");
            for (auto c : took) score[(size_t)c] = new_score[(size_t)c];
        }
        tmp.keep(fresh);
        dropSample();
        fba_belief_destroy(_belief);
        _belief         = fresh;
        _log_likelihood = 0.0; // :394
        ++_chains;
    }
};

} // namespace fba_b200

#endif // FBA_B200_CUDA_MH_HPP
