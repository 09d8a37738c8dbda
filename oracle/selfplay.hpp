// oracle/selfplay.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the self-play driver:
//   synthesis/src/alpha_zero.rs:120-169  gather_experience (worker scheduling + seeding)
//   synthesis/src/alpha_zero.rs:181-209  run_n_games
//   synthesis/src/alpha_zero.rs:211-338  StateInfo, run_game, sample_action, fill_state_info, store_rewards
//   synthesis/src/data.rs:106-194        ReplayBuffer::{new,new_game,add,extend,keep_last_n_games}
// No reference test covers any of this ("parity unpinned" beyond what the MCTS/Connect4 tests pin).
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <thread>
#include <vector>

#include "../include/syn_streams.h"
#include "policies.hpp"

namespace orc {

// data.rs:106-114
struct ReplayBuffer {
    size_t game_id = 0, steps = 0;
    std::vector<size_t> game_ids;
    std::vector<Connect4> games;
    std::vector<std::array<float, 63>> states;
    std::vector<std::array<float, 9>> pis;
    std::vector<std::array<float, 3>> vs;
    void new_game() { game_id += 1; }                 // data.rs:128-130
    size_t curr_steps() const { return vs.size(); }   // data.rs:147-149
    size_t curr_games() const {                        // data.rs:136-140
        size_t n = 0;
        for (size_t i = 0; i < game_ids.size(); ++i)
            if (i == 0 || game_ids[i] != game_ids[i - 1]) ++n;
        return n;
    }
    void add(const Connect4& g, const float pi[9], const float v[3]) { // data.rs:151-158
        game_ids.push_back(game_id);
        steps += 1;
        games.push_back(g);
        std::array<float, 63> s;
        g.features(s.data());
        states.push_back(s);
        pis.push_back({pi[0], pi[1], pi[2], pi[3], pi[4], pi[5], pi[6], pi[7], pi[8]});
        vs.push_back({v[0], v[1], v[2]});
    }
    void extend(ReplayBuffer& o) { // data.rs:160-170
        steps += o.steps;
        size_t start = game_id;
        for (size_t g : o.game_ids) game_ids.push_back(g + start);
        game_id += o.game_id;
        games.insert(games.end(), o.games.begin(), o.games.end());
        states.insert(states.end(), o.states.begin(), o.states.end());
        pis.insert(pis.end(), o.pis.begin(), o.pis.end());
        vs.insert(vs.end(), o.vs.begin(), o.vs.end());
        o.games.clear(); o.states.clear(); o.pis.clear(); o.vs.clear();
    }
    void keep_last_n_games(size_t n) { // data.rs:172-194
        if (game_id <= n) return;
        size_t min_game_id = game_id - n;
        size_t remove = 0;
        for (size_t i = 0; i < game_ids.size(); ++i) {
            if (game_ids[i] >= min_game_id) break;
            remove = i + 1;
        }
        if (remove) {
            game_ids.erase(game_ids.begin(), game_ids.begin() + remove);
            games.erase(games.begin(), games.begin() + remove);
            states.erase(states.begin(), states.begin() + remove);
            pis.erase(pis.begin(), pis.begin() + remove);
            vs.erase(vs.begin(), vs.begin() + remove);
        }
    }
};

struct StateInfo { // alpha_zero.rs:211-227
    size_t turn;
    float t;
    float q[3];
    float z[3];
};

struct GameTrace { // per-ply observables the parity tests compare (not part of the reference)
    std::vector<uint8_t> actions;
    std::vector<uint32_t> tree_nodes;
    std::vector<std::array<float, 9>> child_visits;
};

// alpha_zero.rs:270-294.  Returns -1 where WeightedIndex::new(..).unwrap() would panic.
static inline int sample_action(const syn_rollout_cfg& cfg, MCTS<Connect4>& mcts, const Connect4& game,
                                const float pi[9], StdRng& rng, size_t num_turns) {
    int best = mcts.best_action(cfg.action_selection);
    Outcome solution = mcts.solution(best);
    if (num_turns < cfg.random_actions_until) {
        int acts[9];
        int n = game.actions(acts);
        uint32_t i = rng.gen_range_u8((uint32_t)n);
        return acts[i];
    } else if (num_turns < cfg.sample_actions_until && (!solution.is_some() || !cfg.stop_games_when_solved)) {
        return weighted_index_sample(pi, 9, rng);
    }
    return best;
}

// alpha_zero.rs:229-268 + 296-338.  Returns false if the reference would have panicked.
static inline bool run_game(const syn_rollout_cfg& cfg, Policy<Connect4>* policy, StdRng& rng, ReplayBuffer& buffer,
                            TreeOptions opt, StdRng* noise_rng, StdRng* fpu_rng, Counters* cnt, GameTrace* trace) {
    Connect4 game;
    Outcome solution = Outcome::none();
    float pi[9];
    size_t num_turns = 0;
    std::vector<StateInfo> infos;
    infos.reserve(Connect4::MAX_TURNS);
    while (!solution.is_some()) {
        MCTS<Connect4> mcts(cfg.num_explores + 1, cfg.mcts, policy, game, opt, noise_rng, fpu_rng, cnt);
        mcts.explore_n(cfg.num_explores);
        mcts.finish();
        mcts.target_policy(pi);
        const float zero3[3] = {0.0f, 0.0f, 0.0f};
        buffer.add(game, pi, zero3);
        StateInfo si;
        si.turn = num_turns + 1;
        si.t = 0.0f;
        mcts.target_q(si.q);
        si.z[0] = si.z[1] = si.z[2] = 0.0f;
        infos.push_back(si);

        int action = sample_action(cfg, mcts, game, pi, rng, num_turns);
        if (action < 0) return false;
        if (trace) {
            trace->actions.push_back((uint8_t)action);
            trace->tree_nodes.push_back((uint32_t)mcts.nodes.size());
            std::array<float, 9> cv{};
            const auto& r = mcts.nodes[mcts.root];
            for (uint32_t c = r.first_child; c < r.last_child(); ++c) cv[mcts.nodes[c].action] = mcts.nodes[c].num_visits;
            trace->child_visits.push_back(cv);
        }
        solution = mcts.solution(action);
        bool over = game.step(action);
        if (over) solution = Outcome::from_f32(game.reward(game.player()));
        else if (!cfg.stop_games_when_solved) solution = Outcome::none();
        num_turns += 1;
    }
    // fill_state_info (alpha_zero.rs:296-307)
    Outcome outcome = solution.reversed();
    size_t n = infos.size();
    for (size_t k = n; k-- > 0;) {
        infos[k].z[outcome.index()] = 1.0f;
        infos[k].t = (float)infos[k].turn / (float)n;
        outcome = outcome.reversed();
    }
    // store_rewards (alpha_zero.rs:309-338)
    size_t start = buffer.curr_steps() - n;
    for (size_t k = 0; k < n; ++k) {
        const StateInfo& s = infos[k];
        auto& v = buffer.vs[start + k];
        switch (cfg.value_target_kind) {
        case SYN_VALUE_Q: for (int i = 0; i < 3; ++i) v[i] = s.q[i]; break;
        case SYN_VALUE_Z: for (int i = 0; i < 3; ++i) v[i] = s.z[i]; break;
        case SYN_VALUE_QZ_AVERAGE: {
            float p = cfg.vt_a;
            for (int i = 0; i < 3; ++i) v[i] = s.q[i] * p + s.z[i] * (1.0f - p);
            break;
        }
        default: {
            float p = (1.0f - s.t) * cfg.vt_a + s.t * cfg.vt_b;
            for (int i = 0; i < 3; ++i) v[i] = s.q[i] * (1.0f - p) + s.z[i] * p;
            break;
        }
        }
    }
    return true;
}

// How leaves are evaluated in a gather call.
struct LeafSource {
    const float* weights = nullptr; // Connect4Net blob (NN mode)
    orc_eval_fn callback = nullptr; // overrides weights: outputs supplied from outside
    void* callback_ctx = nullptr;
    bool use_cache = true;          // PolicyWithCache like run_n_games (alpha_zero.rs:196-198)
};

// One game with the engine's per-game streams (include/syn_streams.h).
static inline bool run_game_streams(const syn_rollout_cfg& cfg, const LeafSource& leaf, uint64_t seed, uint64_t g,
                                    ReplayBuffer& buffer, TreeOptions opt, Counters* cnt, GameTrace* trace,
                                    Policy<Connect4>* shared_nn_policy) {
    StdRng rollout_rng = StdRng::seed_from_u64(syn_stream_seed(seed, g, SYN_STREAM_ROLLOUT));
    StdRng action_rng = StdRng::seed_from_u64(syn_stream_seed(seed, g, SYN_STREAM_ACTION));
    StdRng noise_rng = StdRng::seed_from_u64(syn_stream_seed(seed, g, SYN_STREAM_NOISE));
    StdRng fpu_rng = StdRng::seed_from_u64(syn_stream_seed(seed, g, SYN_STREAM_FPU));
    RolloutPolicy<Connect4> rp(&rollout_rng, cnt);
    Policy<Connect4>* p = cfg.leaf_eval_kind == SYN_LEAF_ROLLOUT ? (Policy<Connect4>*)&rp : shared_nn_policy;
    return run_game(cfg, p, action_rng, buffer, opt, &noise_rng, &fpu_rng, cnt, trace);
}

struct GatherResult {
    ReplayBuffer buffer;
    Counters counters;
    uint64_t elapsed_ns = 0;
    uint64_t cache_hits = 0, cache_misses = 0;
    bool ok = true;
};

// Engine-compatible gather: games [first, first+n) with per-game streams, `threads` OS threads
// pulling games from a shared counter; rows are emitted in game order whatever the thread count.
static inline GatherResult gather_streams(const syn_rollout_cfg& cfg, const LeafSource& leaf, uint64_t seed,
                                          uint64_t first, uint32_t n, int threads, TreeOptions opt,
                                          std::vector<GameTrace>* traces) {
    GatherResult res;
    std::vector<ReplayBuffer> per_game(n);
    if (traces) traces->assign(n, GameTrace());
    std::atomic<uint32_t> next(0);
    std::atomic<bool> ok(true);
    if (threads < 1) threads = 1;
    std::vector<Counters> cnts(threads);
    std::vector<uint64_t> hits(threads, 0), misses(threads, 0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        Connect4Net net(leaf.weights, opt.libm);
        CallbackPolicy cb(leaf.callback, leaf.callback_ctx);
        Policy<Connect4>* base = leaf.callback ? (Policy<Connect4>*)&cb : (Policy<Connect4>*)&net;
        PolicyWithCache<Connect4> cached(Connect4::MAX_TURNS * (size_t)n / (size_t)threads + 64, base);
        Policy<Connect4>* nn = leaf.use_cache ? (Policy<Connect4>*)&cached : base;
        for (;;) {
            uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            per_game[i].new_game();
            if (!run_game_streams(cfg, leaf, seed, first + i, per_game[i], opt, &cnts[tid],
                                  traces ? &(*traces)[i] : nullptr, nn))
                ok = false;
        }
        hits[tid] = cached.hits;
        misses[tid] = cached.misses;
    };
    if (threads == 1) work(0);
    else {
        std::vector<std::thread> ts;
        for (int t = 0; t < threads; ++t) ts.emplace_back(work, t);
        for (auto& t : ts) t.join();
    }
    res.elapsed_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    res.buffer.game_id = first; // so that game ids come out as first+1+i (1-based like new_game)
    for (uint32_t i = 0; i < n; ++i) res.buffer.extend(per_game[i]);
    for (int t = 0; t < threads; ++t) {
        res.counters.add(cnts[t]);
        res.cache_hits += hits[t];
        res.cache_misses += misses[t];
    }
    res.ok = ok;
    return res;
}

// The reference's own schedule (alpha_zero.rs:120-209): num_workers+1 threads, games split
// `remaining / workers_left`, worker_seed = seed*(num_workers+1)+w, ONE StdRng per worker for
// action sampling, one private weight copy + memo cache per worker, buffers joined in worker
// order.  NN leaves only (gather_experience requires P: NNPolicy).
static inline GatherResult gather_reference_schedule(const syn_rollout_cfg& cfg, const LeafSource& leaf, size_t num_workers,
                                                     size_t games_per_train, size_t seed, TreeOptions opt) {
    GatherResult res;
    size_t nw = num_workers + 1;
    std::vector<ReplayBuffer> bufs(nw);
    std::vector<Counters> cnts(nw);
    std::vector<uint64_t> hits(nw, 0), misses(nw, 0);
    std::vector<std::thread> ts;
    std::atomic<bool> ok(true);
    size_t to_schedule = games_per_train, left = nw;
    auto t0 = std::chrono::steady_clock::now();
    for (size_t w = 0; w < nw; ++w) {
        size_t num_games = to_schedule / left;
        size_t worker_seed = seed * nw + w;
        ts.emplace_back([&, w, num_games, worker_seed]() {
            StdRng rng = StdRng::seed_from_u64((uint64_t)worker_seed);
            // noise / fpu streams stand in for thread_rng: one per worker
            StdRng noise_rng = StdRng::seed_from_u64(syn_stream_seed(worker_seed, 0, SYN_STREAM_NOISE));
            StdRng fpu_rng = StdRng::seed_from_u64(syn_stream_seed(worker_seed, 0, SYN_STREAM_FPU));
            std::vector<float> private_weights(leaf.weights, leaf.weights + SYN_N_WEIGHTS); // vs.load per worker
            Connect4Net net(private_weights.data(), opt.libm);
            PolicyWithCache<Connect4> cached(Connect4::MAX_TURNS * games_per_train, &net);
            Policy<Connect4>* p = leaf.use_cache ? (Policy<Connect4>*)&cached : (Policy<Connect4>*)&net;
            for (size_t k = 0; k < num_games; ++k) {
                bufs[w].new_game();
                if (!run_game(cfg, p, rng, bufs[w], opt, &noise_rng, &fpu_rng, &cnts[w], nullptr)) ok = false;
            }
            hits[w] = cached.hits;
            misses[w] = cached.misses;
        });
        to_schedule -= num_games;
        left -= 1;
    }
    for (auto& t : ts) t.join();
    res.elapsed_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    for (size_t w = 0; w < nw; ++w) {
        res.buffer.extend(bufs[w]);
        res.counters.add(cnts[w]);
        res.cache_hits += hits[w];
        res.cache_misses += misses[w];
    }
    res.ok = ok;
    return res;
}

} // namespace orc
