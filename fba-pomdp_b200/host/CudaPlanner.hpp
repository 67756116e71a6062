// C++ host adapter for the reference's Planner interface (samkatt/fba-pomdp): POMCP whose simulator
// calls run batched on the GPU.
//
//   CudaBatchedPOUCT : planners::BAPlanner      stands in for planners::RBAPOUCT
//   (src/planners/bayes-adaptive/RBAPOUCT.cpp:67-153; interface src/planners/bayes-adaptive/BAPlanner.hpp:23-40)
//
// The reference runs its `n` simulations one after the other; each one samples a root particle,
// descends the tree with UCB calling BAPOMDP::step, expands one leaf and evaluates it with a
// random-policy rollout. Here the simulations run in WAVES of `wave` simulations:
//   1. the wave's root particles are drawn on the device (fba_belief_sample_batch) — only their
//      indices come to the host, no particle is ever downloaded;
//   2. the tree (host) is descended level by level: every still-active simulation picks its action by
//      UCB — visit counts are incremented at selection time ("virtual visits"), so the simulations of
//      one wave spread over the actions instead of all following the same path — and ONE
//      fba_step_batch call performs the BAPOMDP::step of all of them (KeepCounts, root sampling:
//      the particle keeps its counts, RBAPOUCT.cpp:86-106);
//   3. simulations that open a new leaf stop descending and queue a rollout; ONE fba_rollouts call
//      evaluates all queued leaves (RBAPOUCT::rollout, RBAPOUCT.cpp:295-323);
//   4. returns are backed up along each simulation's path with the reference's update
//      ret = r + discount * delayed (RBAPOUCT.cpp:271-272), running mean per chance node
//      (MCTSTreeNodes.cpp:8-12).
// With wave = 1 this is the reference's algorithm (other random numbers); larger waves trade a little
// search sharpness for a ~wave-fold cut in simulator round trips. The tree, UCB and the choice of
// the final action stay on the host — they are sequential by construction (SURVEY.md §7 hard part 8).
//
// Requires the belief to be one of this repo's CUDA beliefs (the particles must already live on
// the device) and domains whose legal action set is the same in every state (true for all the
// reference's BA domains).
#ifndef FBA_B200_CUDA_PLANNER_HPP
#define FBA_B200_CUDA_PLANNER_HPP

#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <unordered_map>
#include <vector>

#include "CudaBeliefs.hpp"

#include "configurations/Conf.hpp"
#include "environment/History.hpp"
#include "planners/bayes-adaptive/BAPlanner.hpp"
#include "utils/random.hpp"

namespace fba_b200 {

class CudaBatchedPOUCT : public planners::BAPlanner
{
public:
    // c: the reference's own planner flags (-s, --mcts-max-depth, -u, -d, -H); wave: simulations per batch
    explicit CudaBatchedPOUCT(configurations::Conf const& c, int wave = 256, uint64_t seed = 4242) :
            _n(c.planner_conf.mcts_simulation_amount),
            _max_depth(c.planner_conf.mcts_max_depth == -1 ? c.horizon : c.planner_conf.mcts_max_depth),
            _h(c.horizon),
            _u(c.planner_conf.mcts_exploration_const),
            _discount(c.discount),
            _wave(wave)
    {
        // the reference's argument checks (RBAPOUCT.cpp:38-55)
        if (_n < 1) throw "cannot initiate RBAPOUCT with " + std::to_string(_n) + " simulations, must be greater than 0";
        if (_max_depth < 0)
            throw "cannot initiate RBAPOUCT with " + std::to_string(_max_depth) + " max depth, must be greater or equal to 0";
        if (_h <= 0) throw "cannot initiate RBAPOUCT with " + std::to_string(_h) + " horizon, must be greater than 0";
        if (_wave < 1) throw std::string("CudaBatchedPOUCT: wave must be at least 1");
        _rng.mode    = FBA_RNG_PHILOX;
        _rng.words   = nullptr;
        _rng.n_words = _rng.cursor = 0;
        _rng.seed    = seed;
        _rng.offset  = 0;
    }

    Action const* selectAction(BAPOMDP const& simulator, beliefs::BABelief const& belief, History const& history)
        const override
    {
        fba_belief* b  = nullptr;
        fba_ctx* ctx   = nullptr;
        if (auto p = dynamic_cast<CudaParticleBelief const*>(&belief))
        {
            b   = p->handle();
            ctx = p->cuda().ctx();
        } else if (auto r = dynamic_cast<CudaReinvigoratingRejectionSampling const*>(&belief))
        {
            b   = r->handle();
            ctx = r->cuda().ctx();
        } else
            throw std::string("CudaBatchedPOUCT needs one of the fba_b200 CUDA beliefs");

        // the legal actions, from the domain itself (they are the domain's objects, RBAPOUCT.cpp:74)
        std::vector<Action const*> actions;
        simulator.addLegalActions(belief.sample(), &actions);
        int const A = (int)actions.size();
        if (A != simulator.domainSize()->_A)
        {
            for (auto a : actions) simulator.releaseAction(a);
            throw std::string("CudaBatchedPOUCT: state-dependent action sets are not supported");
        }

        int const depth = std::min(_h - (int)history.length(), _max_depth);
        _nodes.clear();
        int const root = newNode(A);

        std::vector<Sim> sims;
        std::vector<int64_t> particle;
        std::vector<int32_t> state, action, new_state, obs, term, start, roll_depth;
        std::vector<int64_t> roll_particle;
        std::vector<double> reward, roll_ret;
        std::vector<int> active, pending;

        for (int done = 0; done < _n; done += _wave)
        {
            int const W = std::min(_wave, _n - done);
            sims.assign(W, Sim());
            particle.resize(W);
            check(ctx, fba_belief_sample_batch(b, &_rng, W, particle.data()), "fba_belief_sample_batch");
            // every root particle starts from its own current domain state: one batched read
            state.resize(W);
            rootStates(ctx, b, particle, &state);
            active.resize(W);
            for (int i = 0; i < W; ++i)
            {
                sims[i].node  = root;
                sims[i].state = state[i];
                active[i]     = i;
            }
            pending.clear();

            for (int d = depth; d > 0 && !active.empty(); --d)
            {
                // UCB with virtual visits (sequential over the wave: each pick sees the earlier ones)
                int const m = (int)active.size();
                std::vector<int64_t> p(m);
                std::vector<int32_t> s(m), a(m);
                for (int k = 0; k < m; ++k)
                {
                    Sim& sim  = sims[active[k]];
                    int const act = pickUCB(sim.node, true);
                    _nodes[sim.node].visits++;
                    _nodes[sim.node].n[act]++;
                    sim.path.push_back({sim.node, act, 0.0});
                    p[k] = particle[active[k]], s[k] = sim.state, a[k] = act;
                }
                new_state.resize(m), obs.resize(m), term.resize(m), reward.resize(m);
                check(ctx,
                      fba_step_batch(b, m, p.data(), s.data(), a.data(), &_rng, new_state.data(), obs.data(),
                                     reward.data(), term.data()),
                      "fba_step_batch");
                std::vector<int> still;
                for (int k = 0; k < m; ++k)
                {
                    Sim& sim              = sims[active[k]];
                    sim.path.back().reward = reward[k];
                    sim.state             = new_state[k];
                    if (term[k]) continue; // terminal: delayed return 0 (RBAPOUCT.cpp:252)
                    auto& children = _nodes[sim.path.back().node].children[sim.path.back().action];
                    auto it        = children.find(obs[k]);
                    if (it != children.end())
                    {
                        sim.node = it->second;
                        if (d - 1 > 0) still.push_back(active[k]);
                    } else
                    { // new leaf: expand, evaluate by rollout (RBAPOUCT.cpp:258-266)
                        int const leaf = newNode(A);
                        _nodes[sim.path.back().node].children[sim.path.back().action][obs[k]] = leaf;
                        sim.rollout_depth = d - 1;
                        pending.push_back(active[k]);
                    }
                }
                active.swap(still);
            }

            // the wave's leaf rollouts, one launch
            if (!pending.empty())
            {
                int const R = (int)pending.size();
                roll_particle.resize(R), start.resize(R), roll_depth.resize(R), roll_ret.resize(R);
                for (int k = 0; k < R; ++k)
                {
                    roll_particle[k] = particle[pending[k]];
                    start[k]         = sims[pending[k]].state;
                    roll_depth[k]    = sims[pending[k]].rollout_depth;
                }
                check(ctx,
                      fba_rollouts(b, R, roll_particle.data(), start.data(), roll_depth.data(), _discount, &_rng,
                                   nullptr, roll_ret.data()),
                      "fba_rollouts");
                for (int k = 0; k < R; ++k) sims[pending[k]].leaf_value = roll_ret[k];
            }

            // back up (RBAPOUCT.cpp:271-272; ChanceNode::addVisit, MCTSTreeNodes.cpp:8-12)
            for (auto& sim : sims)
            {
                double ret = sim.leaf_value;
                for (auto it = sim.path.rbegin(); it != sim.path.rend(); ++it)
                {
                    ret        = it->reward + _discount * ret;
                    Node& node = _nodes[it->node];
                    node.done[it->action]++;
                    node.q[it->action] += (ret - node.q[it->action]) / node.done[it->action];
                }
            }
        }

        int const best = pickUCB(root, false);
        Action const* chosen = nullptr;
        for (auto a : actions)
            if (a->index() == best) chosen = simulator.copyAction(a);
        for (auto a : actions) simulator.releaseAction(a);
        return chosen;
    }

private:
    struct Node
    {
        int visits = 0;
        std::vector<int> n, done;                             // selections (incl. in flight), completed back-ups
        std::vector<double> q;                                // running mean return per action
        std::vector<std::unordered_map<int, int>> children;   // [action][observation] -> node
    };
    struct Step
    {
        int node, action;
        double reward;
    };
    struct Sim
    {
        int node = 0, state = 0, rollout_depth = 0;
        double leaf_value = 0.0;
        std::vector<Step> path;
    };

    int _n, _max_depth, _h;
    double _u, _discount;
    int _wave;
    mutable fba_rng _rng;
    mutable std::vector<Node> _nodes;

    int newNode(int A) const
    {
        Node nd;
        nd.n.assign(A, 0), nd.done.assign(A, 0), nd.q.assign(A, 0.0), nd.children.resize(A);
        _nodes.push_back(std::move(nd));
        return (int)_nodes.size() - 1;
    }

    // argmax_a q(a) [+ u * sqrt(log(1 + m) / n_a), infinite for n_a = 0 (RBAPOUCT.cpp:349-357)];
    // ties broken uniformly (RBAPOUCT.cpp:204)
    int pickUCB(int node_id, bool explore) const
    {
        Node const& node = _nodes[node_id];
        double best      = -std::numeric_limits<double>::max();
        std::vector<int> cand;
        for (int a = 0; a < (int)node.q.size(); ++a)
        {
            double v = node.q[a];
            if (explore)
                v = (node.n[a] == 0) ? std::numeric_limits<double>::max()
                                     : v + _u * std::sqrt(std::log1p((double)node.visits) / node.n[a]);
            if (v > best)
            {
                best = v;
                cand.clear();
            }
            if (v >= best) cand.push_back(a);
        }
        return cand[rnd::slowRandomInt(0, (int)cand.size())];
    }

    static void rootStates(fba_ctx* ctx, fba_belief* b, std::vector<int64_t> const& particle,
                           std::vector<int32_t>* state)
    {
        check(ctx, fba_belief_gather_states(b, (int64_t)particle.size(), particle.data(), state->data()),
              "fba_belief_gather_states");
    }
};

// POMCP with the search tree itself on the device (fba_tree_*): every simulation — root particle,
// UCB descent, leaf expansion, rollout, back-up — runs inside one kernel, `wave` simulations at a
// time sharing the tree through atomics. The host only sees the root's action values. Stands in for
// planners::RBAPOUCT like CudaBatchedPOUCT, with no per-level round trip between host and device.
class CudaTreePOUCT : public planners::BAPlanner
{
public:
    explicit CudaTreePOUCT(configurations::Conf const& c, int wave = 256, uint64_t seed = 777) :
            _n(c.planner_conf.mcts_simulation_amount),
            _max_depth(c.planner_conf.mcts_max_depth == -1 ? c.horizon : c.planner_conf.mcts_max_depth),
            _h(c.horizon),
            _u(c.planner_conf.mcts_exploration_const),
            _discount(c.discount),
            _wave(wave)
    {
        // the reference's argument checks (RBAPOUCT.cpp:38-55)
        if (_n < 1) throw "cannot initiate RBAPOUCT with " + std::to_string(_n) + " simulations, must be greater than 0";
        if (_max_depth < 0)
            throw "cannot initiate RBAPOUCT with " + std::to_string(_max_depth) + " max depth, must be greater or equal to 0";
        if (_h <= 0) throw "cannot initiate RBAPOUCT with " + std::to_string(_h) + " horizon, must be greater than 0";
        if (_wave < 1) throw std::string("CudaTreePOUCT: wave must be at least 1");
        _rng.mode    = FBA_RNG_PHILOX;
        _rng.words   = nullptr;
        _rng.n_words = _rng.cursor = 0;
        _rng.seed    = seed;
        _rng.offset  = 0;
    }

    Action const* selectAction(BAPOMDP const& simulator, beliefs::BABelief const& belief, History const& history)
        const override
    {
        fba_belief* b               = nullptr;
        CudaSimulator const* cuda = nullptr;
        if (auto p = dynamic_cast<CudaParticleBelief const*>(&belief)) b = p->handle(), cuda = &p->cuda();
        else if (auto r = dynamic_cast<CudaReinvigoratingRejectionSampling const*>(&belief))
            b = r->handle(), cuda = &r->cuda();
        else
            throw std::string("CudaTreePOUCT needs one of the fba_b200 CUDA beliefs");
        fba_ctx* ctx   = cuda->ctx();
        fba_tree* tree = cuda->tree(_n, std::max(1, std::max(_max_depth, _h)));
        int const depth = std::min(_h - (int)history.length(), _max_depth);
        int32_t best    = 0;
        check(ctx, fba_tree_search(tree, b, _n, depth, _u, _discount, _wave, &_rng, &best, nullptr, nullptr),
              "fba_tree_search");
        // the domain's own action object for that index (RBAPOUCT.cpp:74)
        std::vector<Action const*> actions;
        simulator.addLegalActions(belief.sample(), &actions);
        Action const* chosen = nullptr;
        for (auto a : actions)
            if (a->index() == best) chosen = simulator.copyAction(a);
        for (auto a : actions) simulator.releaseAction(a);
        if (!chosen) throw std::string("CudaTreePOUCT: state-dependent action sets are not supported");
        return chosen;
    }

private:
    int _n, _max_depth, _h;
    double _u, _discount;
    int _wave;
    mutable fba_rng _rng;
};

} // namespace fba_b200

#endif // FBA_B200_CUDA_PLANNER_HPP
