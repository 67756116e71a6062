// Drop-ins for the reference's COMPOSITE structure beliefs, behind the same BABelief
// interface — the classes whose update is a fixed sequence of particle-filter primitives on two or three
// filters, each of which is one call into libfba_b200.so here:
//
//   CudaCheatingReinvigoration      stands in for beliefs::bayes_adaptive::prototypes::CheatingReinvigoration
//                                   (src/beliefs/bayes-adaptive/prototypes/CheatingReinvigoration.{hpp,cpp})
//   CudaStructureIncubatorSampling  stands in for beliefs::bayes_adaptive::factored::StructureIncubatorSampling
//                                   (src/beliefs/bayes-adaptive/factored/StructureIncubatorSampling.{hpp,cpp})
//   CudaNestedBelief                stands in for beliefs::bayes_adaptive::NestedBelief
//                                   (src/beliefs/bayes-adaptive/NestedBelief.{hpp,cpp})
//
// What runs where: initiate samples the reference's own prior on the host (sampleStartState,
// sampleCorrectGraphState, sampleFullyConnectedState) and uploads the particles; domain start states for
// resetDomainStateDistribution come from the reference's own domain; everything per update — rejection
// sampling, importance-sampling update + resample, breeding (mutate + marginalizeOut), cheating copies,
// promotion of heavy shadow particles — runs on the GPU filters. Same constructor checks and error strings
// as the reference. Registered by a maintainer in factory::makeBABelief (BABelief.cpp:56-71) next to
// "incubator" and "cheating-reinvigoration" (INTEGRATION.md).
#ifndef FBA_B200_CUDA_STRUCTURE_BELIEFS_HPP
#define FBA_B200_CUDA_STRUCTURE_BELIEFS_HPP

#include <cmath>
#include <functional>
#include <string>
#include <vector>

#include "CudaBeliefs.hpp"

namespace fba_b200 {

// Several particle filters over ONE simulator / model: the shared plumbing of the composite beliefs.
class CudaFilterSet : public beliefs::BABelief
{
public:
    CudaFilterSet(size_t size, uint64_t seed, int device) : _size(size), _device(device)
    {
        _rng.mode    = FBA_RNG_PHILOX;
        _rng.words   = nullptr;
        _rng.n_words = _rng.cursor = 0;
        _rng.seed    = seed;
        _rng.offset  = 0;
    }
    ~CudaFilterSet() override { release(); }

    void free(POMDP const& /*d*/) override { release(); }

    // Belief::sample on filter 0 (the belief proper in every composite class)
    State const* sample() const override
    {
        int64_t i = 0;
        check(_cuda->ctx(), fba_belief_sample(_f[0], &_rng, &i), "fba_belief_sample");
        std::vector<float> counts((size_t)fba_belief_stride(_f[0]));
        int32_t state = 0, sid = 0;
        check(_cuda->ctx(), fba_belief_download(_f[0], i, 1, &state, &sid, counts.data(), nullptr), "fba_belief_download");
        dropSample();
        _sample = _cuda->materialise(sid, state, counts);
        return _sample;
    }

    // BAPOMDP::resetDomainState on every particle of every filter, in filter order, counts / weights kept
    void resetDomainStateDistribution(BAPOMDP const& bapomdp) override
    {
        std::vector<int32_t> state(_size);
        for (auto f : _reset_order)
        {
            for (size_t i = 0; i < _size; ++i)
            {
                auto s   = bapomdp.sampleDomainState();
                state[i] = s->index();
                bapomdp.releaseDomainState(s);
            }
            check(_cuda->ctx(), fba_belief_upload(_f[(size_t)f], 0, (int64_t)_size, state.data(), nullptr, nullptr, nullptr),
                  "fba_belief_upload");
        }
    }

    fba_belief* handle(size_t k = 0) const { return _f[k]; }
    CudaSimulator const& cuda() const { return *_cuda; }

protected:
    size_t _size;
    int _device;
    mutable fba_rng _rng;
    std::unique_ptr<CudaSimulator> _cuda;
    std::vector<fba_belief*> _f;
    std::vector<int> _reset_order;
    mutable BAState const* _sample = nullptr;

    struct HostParticles
    {
        std::vector<int32_t> state, sid;
        std::vector<std::vector<float>> blocks;
        size_t stride = 0;
    };

    // `_size` particles from a host-side sampler of the reference (prior x domain start state)
    HostParticles sampleOnHost(POMDP const& d, std::function<BAState const*()> const& draw) const
    {
        HostParticles p;
        p.state.resize(_size), p.sid.resize(_size), p.blocks.resize(_size);
        for (size_t i = 0; i < _size; ++i)
        {
            BAState const* s = draw();
            p.state[i]       = s->_domain_state->index();
            p.sid[i]         = _cuda->describe(s, &p.blocks[i]);
            p.stride         = std::max(p.stride, p.blocks[i].size());
            d.releaseState(s);
        }
        return p;
    }

    fba_belief* makeFilter(HostParticles const& p, size_t stride, bool weighted)
    {
        fba_belief* b = nullptr;
        check(_cuda->ctx(), fba_belief_create(_cuda->ctx(), _cuda->model(), (int64_t)_size, (int64_t)stride, weighted ? 1 : 0, &b),
              "fba_belief_create");
        _f.push_back(b);
        size_t const st = (size_t)fba_belief_stride(b);
        std::vector<float> flat(_size * st, 0.0f);
        for (size_t i = 0; i < _size; ++i) std::copy(p.blocks[i].begin(), p.blocks[i].end(), flat.begin() + i * st);
        std::vector<double> w(weighted ? _size : 0, 1.0 / (double)_size);
        check(_cuda->ctx(),
              fba_belief_upload(b, 0, (int64_t)_size, p.state.data(), p.sid.data(), flat.data(), weighted ? w.data() : nullptr),
              "fba_belief_upload");
        return b;
    }

    // cells the fully connected structure needs: bounds everything breeding / cheating can put in a block
    size_t fullyConnectedStride(POMDP const& d) const
    {
        auto const& fbapomdp = dynamic_cast<::bayes_adaptive::factored::FBAPOMDP const&>(d);
        auto p               = static_cast<BAState const*>(fbapomdp.sampleFullyConnectedState());
        std::vector<float> block;
        _cuda->describe(p, &block);
        d.releaseState(p);
        return block.size();
    }

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
        for (auto b : _f) fba_belief_destroy(b);
        _f.clear();
        _cuda.reset();
    }
};

class CudaCheatingReinvigoration : public CudaFilterSet
{
public:
    CudaCheatingReinvigoration(size_t size, size_t cheat_amount, double resample_threshold, uint64_t seed = 42,
                               int device = 0) :
            CudaFilterSet(size, seed, device), _cheat_amount(cheat_amount), _resample_threshold(resample_threshold)
    {
        if (_size < 1 || _cheat_amount < 1) // as CheatingReinvigoration.cpp:34-38
            throw "CheatingReinvigoration::cannot initiate belief of size < 1 (" + std::to_string(_size)
                + "), or resample size of < 1 (" + std::to_string(_cheat_amount) + ")";
        if (_resample_threshold >= 0) // :40-44
            throw "CheatingReinvigoration::cannot initiate with resample_threshold >= 0 (is:"
                + std::to_string(_resample_threshold) + ")";
    }

    void initiate(POMDP const& d) override
    {
        auto const& fbapomdp = dynamic_cast<::bayes_adaptive::factored::FBAPOMDP const&>(d);
        _cuda.reset(new CudaSimulator(fbapomdp, _device, 1 << 16));
        // :68-93: _size x sampleCorrectGraphState, then _size x sampleStartState with weight 1 / _size
        auto correct = sampleOnHost(d, [&] { return static_cast<BAState const*>(fbapomdp.sampleCorrectGraphState()); });
        auto belief  = sampleOnHost(d, [&] { return static_cast<BAState const*>(d.sampleStartState()); });
        size_t const stride = std::max(std::max(correct.stride, belief.stride), fullyConnectedStride(d));
        makeFilter(belief, stride, true);   // _f[0] = _belief
        makeFilter(correct, stride, false); // _f[1] = _correct_structured_belief
        _reset_order = {1, 0};              // :50-66
        _likelihood  = 1;
        _cheats      = 0;
    }

    void updateEstimation(Action const* a, Observation const* o, POMDP const& /*d*/) override
    { // :107-134
        int64_t attempts = 0;
        check(_cuda->ctx(), fba_belief_reject_sample(_f[1], a->index(), o->index(), &_rng, &attempts),
              "fba_belief_reject_sample");
        double l = 0.0;
        check(_cuda->ctx(), fba_belief_update_estimation(_f[0], a->index(), o->index(), &_rng, &l),
              "fba_belief_update_estimation");
        _likelihood *= l;
        if (std::log(_likelihood) < _resample_threshold)
        {
            check(_cuda->ctx(), fba_belief_cheat(_f[0], _f[1], (int64_t)_cheat_amount, &_rng), "fba_belief_cheat"); // :136-147
            _likelihood = 1;
            ++_cheats;
        }
    }

    size_t cheats() const { return _cheats; }

private:
    size_t _cheat_amount;
    double _resample_threshold;
    double _likelihood = 1;
    size_t _cheats     = 0;
};

class CudaStructureIncubatorSampling : public CudaFilterSet
{
public:
    CudaStructureIncubatorSampling(size_t size, size_t reinvigor_amount, double threshold, int mutate_kind,
                                   uint64_t seed = 42, int device = 0) :
            CudaFilterSet(size, seed, device),
            _shadow_reinvigor_amount(reinvigor_amount),
            _real_reinvigor_threshold(threshold),
            _mutate(mutate_kind)
    {
        if (size < 1 || _shadow_reinvigor_amount < 1) // as StructureIncubatorSampling.cpp:29-34
            throw "StructureIncubatorSampling::Cannot initiate Incubator belief update with size < 1 ("
                + std::to_string(_size) + ") or resample size < 1 (" + std::to_string(_shadow_reinvigor_amount) + ")";
        if (_real_reinvigor_threshold <= 0 || _real_reinvigor_threshold > 1) // :36-40
            throw "StructureIncubatorSampling::must initiate with 1 < threshold <= 0 (is:"
                + std::to_string(_real_reinvigor_threshold) + ")";
        if (_shadow_reinvigor_amount >= size) // WeightedFilter::leastLikely asserts n < size (WeightedFilter.cpp:208)
            throw std::string("StructureIncubatorSampling::resample size must be smaller than the belief");
    }

    void initiate(POMDP const& d) override
    {
        auto const& fbapomdp = dynamic_cast<::bayes_adaptive::factored::FBAPOMDP const&>(d);
        _cuda.reset(new CudaSimulator(fbapomdp, _device, 1 << 16));
        // :65-73: _size x sampleStartState, _size x sampleFullyConnectedState
        auto belief = sampleOnHost(d, [&] { return static_cast<BAState const*>(d.sampleStartState()); });
        auto fc     = sampleOnHost(d, [&] { return static_cast<BAState const*>(fbapomdp.sampleFullyConnectedState()); });
        size_t const stride = std::max(belief.stride, fc.stride);
        makeFilter(belief, stride, false); // _f[0] = _belief
        makeFilter(fc, stride, false);     // _f[1] = _fully_connected_belief
        makeFilter(belief, stride, true);  // _f[2] = _shadow_belief: placeholders, every slot is bred over below
        // :74-80: the shadow belief starts as _size bred particles of weight 1 / _size
        std::vector<int64_t> all(_size);
        for (size_t i = 0; i < _size; ++i) all[i] = (int64_t)i;
        check(_cuda->ctx(), fba_belief_breed_into(_f[2], all.data(), (int64_t)_size, _f[0], _f[1], _mutate, &_rng),
              "fba_belief_breed_into");
        std::vector<double> w(_size, 1.0 / (double)_size);
        check(_cuda->ctx(), fba_belief_upload(_f[2], 0, (int64_t)_size, nullptr, nullptr, nullptr, w.data()),
              "fba_belief_upload");
        _reset_order = {0, 1, 2}; // :46-61
        _promoted    = 0;
    }

    void updateEstimation(Action const* a, Observation const* o, POMDP const& /*d*/) override
    { // :107-137
        fba_ctx* ctx = _cuda->ctx();
        int64_t n    = 0;
        check(ctx, fba_belief_promote(_f[2], _f[0], _real_reinvigor_threshold, &_rng, &n), "fba_belief_promote"); // :155-187
        _promoted += (size_t)n;
        std::vector<int64_t> least(_shadow_reinvigor_amount); // :139-153
        check(ctx, fba_belief_least_likely(_f[2], (int64_t)least.size(), least.data()), "fba_belief_least_likely");
        check(ctx, fba_belief_breed_into(_f[2], least.data(), (int64_t)least.size(), _f[0], _f[1], _mutate, &_rng),
              "fba_belief_breed_into");
        int64_t attempts = 0;
        check(ctx, fba_belief_reject_sample(_f[0], a->index(), o->index(), &_rng, &attempts), "fba_belief_reject_sample");
        check(ctx, fba_belief_reject_sample(_f[1], a->index(), o->index(), &_rng, &attempts), "fba_belief_reject_sample");
        check(ctx, fba_belief_update_estimation(_f[2], a->index(), o->index(), &_rng, nullptr),
              "fba_belief_update_estimation");
    }

    size_t promoted() const { return _promoted; }

private:
    size_t _shadow_reinvigor_amount;
    double _real_reinvigor_threshold;
    int _mutate;
    size_t _promoted = 0;
};

// beliefs::bayes_adaptive::NestedBelief (src/beliefs/bayes-adaptive/NestedBelief.{hpp,cpp}) over fba_nested_*:
// the top filter's count blocks come from the reference's own prior on the host (top_filter_size x
// sampleStartState), the bottom filters' domain states from the reference's own domain; the update — per top
// particle a rejection-sampling loop over its bottom filter that raises its counts by 1 / bottom size per
// accepted state — runs on the GPU, one thread per top particle. Tabular and factored simulators.
class CudaNestedBelief : public beliefs::BABelief
{
public:
    CudaNestedBelief(size_t top_filter_size, size_t bottom_filter_size, uint64_t seed = 42, int device = 0) :
            _top(top_filter_size), _bottom(bottom_filter_size), _device(device)
    {
        if (top_filter_size < 1 || bottom_filter_size < 1) // as NestedBelief.cpp:19-26
            throw "NestedBelief: cannot initiate with filter size < 1 (top: " + std::to_string(_top)
                + ", bottom: " + std::to_string(_bottom) + ")";
        _rng.mode    = FBA_RNG_PHILOX;
        _rng.words   = nullptr;
        _rng.n_words = _rng.cursor = 0;
        _rng.seed    = seed;
        _rng.offset  = 0;
    }
    ~CudaNestedBelief() override { release(); }

    void initiate(POMDP const& d) override
    {
        auto const& bapomdp = dynamic_cast<BAPOMDP const&>(d);
        _cuda.reset(new CudaSimulator(bapomdp, _device, 4096, 0));
        // :63-90: per top particle a prior sample (its domain state is dropped) and `bottom` domain start states
        std::vector<int32_t> sid(_top), states(_top * _bottom);
        std::vector<std::vector<float>> blocks(_top);
        size_t stride = 0;
        for (size_t i = 0; i < _top; ++i)
        {
            for (size_t j = 0; j < _bottom; ++j) states[i * _bottom + j] = startState(bapomdp);
            auto p = static_cast<BAState const*>(bapomdp.sampleStartState());
            sid[i] = _cuda->describe(p, &blocks[i]);
            stride = std::max(stride, blocks[i].size());
            d.releaseState(p);
        }
        check(_cuda->ctx(),
              fba_nested_create(_cuda->ctx(), _cuda->model(), (int64_t)_top, (int64_t)_bottom, (int64_t)stride, &_nested),
              "fba_nested_create");
        fba_belief* top = fba_nested_top(_nested);
        size_t const st = (size_t)fba_belief_stride(top);
        std::vector<float> flat(_top * st, 0.0f);
        for (size_t i = 0; i < _top; ++i) std::copy(blocks[i].begin(), blocks[i].end(), flat.begin() + i * st);
        std::vector<int32_t> zeros(_top, 0);
        std::vector<double> w(_top, 1.0 / (double)_top);
        check(_cuda->ctx(), fba_belief_upload(top, 0, (int64_t)_top, zeros.data(), sid.data(), flat.data(), w.data()),
              "fba_belief_upload");
        check(_cuda->ctx(), fba_nested_upload_states(_nested, 0, (int64_t)_top, states.data()), "fba_nested_upload_states");
    }

    void free(POMDP const& /*d*/) override { release(); }

    // :117-127: the drawn top particle's counts with a state of its bottom filter as domain state
    State const* sample() const override
    {
        int64_t i     = 0;
        int32_t state = 0;
        check(_cuda->ctx(), fba_nested_sample(_nested, &_rng, &i, &state), "fba_nested_sample");
        fba_belief* top = fba_nested_top(_nested);
        std::vector<float> counts((size_t)fba_belief_stride(top));
        int32_t sid = 0;
        check(_cuda->ctx(), fba_belief_download(top, i, 1, nullptr, &sid, counts.data(), nullptr), "fba_belief_download");
        dropSample();
        _sample = _cuda->materialise(sid, state, counts);
        return _sample;
    }

    void updateEstimation(Action const* a, Observation const* o, POMDP const& /*d*/) override
    { // :129-193
        check(_cuda->ctx(), fba_nested_update(_nested, a->index(), o->index(), &_rng, (int64_t)1 << 40, nullptr),
              "fba_nested_update");
    }

    void resetDomainStateDistribution(BAPOMDP const& bapomdp) override
    { // :33-61: fresh start states from the reference's own domain in every bottom filter
        std::vector<int32_t> states(_top * _bottom);
        for (auto& s : states) s = startState(bapomdp);
        check(_cuda->ctx(), fba_nested_upload_states(_nested, 0, (int64_t)_top, states.data()), "fba_nested_upload_states");
    }

    fba_nested* handle() const { return _nested; }

private:
    size_t _top, _bottom;
    int _device;
    mutable fba_rng _rng;
    std::unique_ptr<CudaSimulator> _cuda;
    fba_nested* _nested            = nullptr;
    mutable BAState const* _sample = nullptr;

    static int32_t startState(BAPOMDP const& bapomdp)
    {
        auto s          = bapomdp.sampleDomainState();
        int32_t const i = s->index();
        bapomdp.releaseDomainState(s);
        return i;
    }
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
        fba_nested_destroy(_nested);
        _nested = nullptr;
        _cuda.reset();
    }
};

} // namespace fba_b200

#endif // FBA_B200_CUDA_STRUCTURE_BELIEFS_HPP
