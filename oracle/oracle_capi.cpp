// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// extern "C" surface of the CPU oracle so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg can drive it through ctypes.  Nothing under
// synthesis_b200/ may load this library.
#include <cstdio>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "selfplay.hpp"

using namespace orc;

#define ORC_FLAG_LEGACY 1u
#define ORC_FLAG_LIBM 2u
#define ORC_FLAG_NO_CACHE 4u

static TreeOptions opts_from(uint32_t flags) {
    TreeOptions o;
    o.legacy = (flags & ORC_FLAG_LEGACY) != 0;
    o.libm = (flags & ORC_FLAG_LIBM) != 0;
    return o;
}

static void fill_stats(syn_stats* s, const Counters& c, uint64_t rows, uint64_t games, uint64_t ns) {
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    s->explores = c.explores; s->leaf_evals = c.leaf_evals; s->rows = rows; s->games = games; s->trees = c.trees;
    s->nodes = c.nodes; s->select_levels = c.select_levels; s->children_scanned = c.children_scanned;
    s->expansions = c.expansions; s->children_created = c.children_created; s->backprop_levels = c.backprop_levels;
    s->rollout_plies = c.rollout_plies; s->device_ns = ns;
}

static int copy_experience(const ReplayBuffer& b, syn_experience* out) {
    size_t n = b.curr_steps();
    out->len = n;
    out->games = b.curr_games();
    if (n > out->capacity) return SYN_ERR_CAPACITY;
    for (size_t i = 0; i < n; ++i) {
        if (out->game_ids) out->game_ids[i] = b.game_ids[i];
        if (out->my_bb) out->my_bb[i] = b.games[i].my_bb;
        if (out->op_bb) out->op_bb[i] = b.games[i].op_bb;
        if (out->height) std::memcpy(out->height + 9 * i, b.games[i].height, 9);
        if (out->player) out->player[i] = b.games[i].player_;
        if (out->states) std::memcpy(out->states + 63 * i, b.states[i].data(), 63 * 4);
        if (out->pis) std::memcpy(out->pis + 9 * i, b.pis[i].data(), 9 * 4);
        if (out->vs) std::memcpy(out->vs + 3 * i, b.vs[i].data(), 3 * 4);
    }
    return SYN_OK;
}

extern "C" {

int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

float orc_expf(float x) { return syn_expf(x); }
float orc_logf(float x) { return syn_logf(x); }

// ---- random streams
void orc_chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) { chacha_block(key, counter, rounds, out); }
void orc_seed_key(uint64_t seed, uint32_t key[8]) {
    StdRng r = StdRng::seed_from_u64(seed);
    std::memcpy(key, r.key, 32);
}
void orc_stdrng_words(uint64_t seed, uint32_t n, uint32_t* out) {
    StdRng r = StdRng::seed_from_u64(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = r.next_u32();
}
void orc_gen_range_u8(uint64_t seed, uint32_t n, uint32_t count, uint8_t* out) {
    StdRng r = StdRng::seed_from_u64(seed);
    for (uint32_t i = 0; i < count; ++i) out[i] = (uint8_t)r.gen_range_u8(n);
}
void orc_weighted_index(uint64_t seed, const float* w, int n, uint32_t count, int32_t* out) {
    StdRng r = StdRng::seed_from_u64(seed);
    for (uint32_t i = 0; i < count; ++i) out[i] = weighted_index_sample(w, n, r);
}
void orc_dirichlet(uint64_t seed, float alpha, int k, uint32_t count, float* out) {
    StdRng r = StdRng::seed_from_u64(seed);
    for (uint32_t i = 0; i < count; ++i) syn_dirichlet(r, alpha, k, out + (size_t)i * k);
}
void orc_normal(uint64_t seed, float mean, float std_, uint32_t count, float* out) {
    StdRng r = StdRng::seed_from_u64(seed);
    for (uint32_t i = 0; i < count; ++i) out[i] = syn_normal(r, mean, std_);
}
uint64_t orc_stream_seed(uint64_t seed, uint64_t g, unsigned k) { return syn_stream_seed(seed, g, k); }

// ---- outcome lattice: packed Option<Outcome> as in synthesis_b200.h
static Outcome unpack(uint8_t p) {
    if (p == 0) return Outcome::none();
    Outcome o;
    o.kind = (uint8_t)(p >> 6);
    o.turns = p & 63u;
    return o;
}
int orc_outcome_cmp(uint8_t a, uint8_t b) { return outcome_cmp(unpack(a), unpack(b)); }
uint8_t orc_outcome_reversed(uint8_t a) { return unpack(a).reversed().packed(); }
uint8_t orc_outcome_from_f32(float v) { return Outcome::from_f32(v).packed(); }
float orc_outcome_value(uint8_t a) { return unpack(a).value(); }

// ---- Connect4
int orc_c4_won(uint64_t bb) { return Connect4::won(bb) ? 1 : 0; }
// plays moves[0..n); status bit0 = is_over, bit1 = previous mover won; returns 255 on an illegal move
int orc_c4_play(const uint8_t* moves, uint32_t n, uint64_t* my_bb, uint64_t* op_bb, uint8_t* height9, uint32_t* legal_mask,
                uint8_t* status, float* features63, float* reward_to_move, uint8_t* step_over /*[n] or NULL*/) {
    Connect4 g;
    for (uint32_t i = 0; i < n; ++i) {
        if (moves[i] >= 9 || g.height[moves[i]] >= 7) return 255;
        bool over = g.step(moves[i]);
        if (step_over) step_over[i] = over ? 1 : 0;
    }
    if (my_bb) *my_bb = g.my_bb;
    if (op_bb) *op_bb = g.op_bb;
    if (height9) std::memcpy(height9, g.height, 9);
    if (legal_mask) {
        int acts[9];
        int na = g.actions(acts);
        uint32_t m = 0;
        for (int i = 0; i < na; ++i) m |= 1u << acts[i];
        *legal_mask = m;
    }
    if (status) *status = (uint8_t)((g.is_over() ? 1 : 0) | (g.winner() >= 0 ? 2 : 0));
    if (features63) g.features(features63);
    if (reward_to_move) *reward_to_move = g.reward(g.player());
    return 0;
}
int orc_c4_player(uint64_t my_bb, uint64_t op_bb) { return Connect4::from_bitboards(my_bb, op_bb).player(); }

// ---- the reference's MCTS known-answer tests on TicTacToe (mcts.rs:691-831)
// which: 0 = test_solve_win, 1 = test_solve_loss, 2 = test_solve_draw
int orc_ttt_kat(int which, uint32_t flags, uint32_t* nodes_len, int* best_action_q, uint8_t child_solution[9], uint8_t* root_solution) {
    StdRng rng = StdRng::seed_from_u64(0);
    RolloutPolicy<TicTacToe> policy(&rng);
    TicTacToe game;
    if (which == 0) { game.step(0); game.step(2); }
    else if (which == 1) { game.step(0); game.step(2); game.step(6); }
    else { game.step(0); game.step(4); }
    syn_mcts_cfg cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.exploration_kind = SYN_EXPLORATION_POLYNOMIAL_UCT;
    cfg.c = 2.0f;
    cfg.solve = 1; cfg.correct_values_on_solve = 1; cfg.select_solved_nodes = 1; cfg.auto_extend = 1;
    cfg.fpu_kind = SYN_FPU_CONST;
    cfg.fpu_a = std::numeric_limits<float>::infinity();
    cfg.noise_kind = SYN_NOISE_NONE;
    MCTS<TicTacToe> mcts(1601, cfg, &policy, game, opts_from(flags));
    size_t guard = 0;
    while (!mcts.nodes[mcts.root].solution.is_some() && guard++ < 1000000) mcts.explore();
    *nodes_len = (uint32_t)mcts.nodes.size();
    *best_action_q = mcts.best_action(SYN_ACTION_Q);
    for (int a = 0; a < 9; ++a) child_solution[a] = mcts.solution(a).packed();
    *root_solution = mcts.nodes[mcts.root].solution.packed();
    return 0;
}

// ---- Connect4Net forward (study-connect4/src/policies.rs:47-59)
// syn_logf (kind 0) / syn_expf (kind 1) over an array: lets tests/test_shared_math_gap.py compare the shared functions with
// the correctly rounded results pointwise.
int orc_detmath(int kind, const float* x, float* y, size_t n) {
    for (size_t i = 0; i < n; ++i) y[i] = kind == 0 ? syn_logf(x[i]) : syn_expf(x[i]);
    return 0;
}

int orc_mlp_eval(const float* weights, const uint64_t* my_bb, const uint64_t* op_bb, uint32_t n, uint32_t flags, float* logits,
                 float* probs) {
    Connect4Net net(weights, (flags & ORC_FLAG_LIBM) != 0);
    for (uint32_t i = 0; i < n; ++i) net.eval(Connect4::from_bitboards(my_bb[i], op_bb[i]), logits + 9 * (size_t)i, probs + 3 * (size_t)i);
    return 0;
}

// ---- one tree from an arbitrary position: what syn_engine_search returns
int orc_search(const syn_rollout_cfg* cfg, uint32_t tree_kind, uint64_t my_bb, uint64_t op_bb, uint64_t rollout_seed,
               const float* weights, orc_eval_fn callback, void* ctx, uint32_t flags, float* child_visits9,
               uint8_t* child_solution9, float* root_q3, uint8_t* root_solution, uint8_t* best_action, uint32_t* num_nodes,
               syn_stats* stats) {
    TreeOptions opt = opts_from(flags);
    Connect4 game = Connect4::from_bitboards(my_bb, op_bb);
    Counters cnt;
    StdRng rng = StdRng::seed_from_u64(rollout_seed);
    // noise/fpu streams for a bare search: derived from the same seed, streams 2 and 3
    StdRng noise_rng = StdRng::seed_from_u64(rollout_seed ^ (1ull << 63));
    StdRng fpu_rng = StdRng::seed_from_u64((rollout_seed ^ (1ull << 63)) + 1);
    RolloutPolicy<Connect4> rp(&rng, &cnt);
    Connect4Net net(weights, opt.libm);
    CallbackPolicy cb(callback, ctx);
    Policy<Connect4>* p = cfg->leaf_eval_kind == SYN_LEAF_ROLLOUT ? (Policy<Connect4>*)&rp
                          : (callback ? (Policy<Connect4>*)&cb : (Policy<Connect4>*)&net);
    if (child_visits9) std::memset(child_visits9, 0, 36);
    if (child_solution9) std::memset(child_solution9, 0, 9);
    if (tree_kind == SYN_TREE_MCTS) {
        MCTS<Connect4> m(cfg->num_explores + 1, cfg->mcts, p, game, opt, &noise_rng, &fpu_rng, &cnt);
        m.explore_n(cfg->num_explores);
        m.finish();
        const auto& r = m.nodes[m.root];
        for (uint32_t c = r.first_child; c < r.last_child(); ++c) {
            if (child_visits9) child_visits9[m.nodes[c].action] = m.nodes[c].num_visits;
            if (child_solution9) child_solution9[m.nodes[c].action] = m.nodes[c].solution.packed();
        }
        if (root_q3) m.target_q(root_q3);
        if (root_solution) *root_solution = r.solution.packed();
        if (best_action) *best_action = (uint8_t)m.best_action(cfg->action_selection);
        if (num_nodes) *num_nodes = (uint32_t)m.nodes.size();
    } else {
        FrozenMCTS<Connect4> m(cfg->num_explores + 1, cfg->mcts, p, game, opt, &cnt);
        m.explore_n(cfg->num_explores);
        m.finish();
        if (m.unsupported) return SYN_ERR_UNSUPPORTED;
        const auto& r = m.nodes[m.root];
        for (uint32_t c = r.first_child; c < r.last_child(); ++c) {
            if (child_visits9) child_visits9[m.nodes[c].action] = m.nodes[c].num_visits;
            if (child_solution9) child_solution9[m.nodes[c].action] = m.nodes[c].solution.packed();
        }
        if (root_q3) { // FrozenMCTS keeps a scalar: report [cum_value, 0, num_visits]
            root_q3[0] = r.cum_value; root_q3[1] = 0.0f; root_q3[2] = r.num_visits;
        }
        if (root_solution) *root_solution = r.solution.packed();
        if (best_action) *best_action = (uint8_t)m.best_action(cfg->action_selection);
        if (num_nodes) *num_nodes = (uint32_t)m.nodes.size();
    }
    fill_stats(stats, cnt, 0, 0, 0);
    return 0;
}

// ---- one evaluation match (evaluator.rs:129-228): players[0] moves first; every move is
// MCTS::exploit (mcts.rs:111-121) or FrozenMCTS::exploit (evaluator.rs:308-318) of the mover's
// player on a clone of the game; all RolloutPolicy draws of the match come from ONE
// StdRng::seed_from_u64(seed) (evaluator.rs:172-173, 207-208).  Returns game.reward(first_player).
// Trace per move: the column played, nodes.len() and the root's child visit counts by column.
// weights2 (optional): players[1]'s network when two different networks meet (eval_against_old, evaluator.rs:131-161);
// mover_out (optional): set to the index of the player whose tree is being searched before every move, so that a
// callback can tell whose leaf it is asked about.
int orc_match2(const syn_player_cfg* players, uint64_t seed, const uint32_t* explores2, const float* weights, const float* weights2,
               uint32_t* mover_out, orc_eval_fn callback, void* ctx, uint32_t flags, float* result, uint8_t* n_moves, uint8_t* moves63,
               uint32_t* tree_nodes63, float* child_visits63x9, syn_stats* stats) {
    TreeOptions opt = opts_from(flags);
    Counters cnt;
    StdRng rng = StdRng::seed_from_u64(seed);
    StdRng noise_rng = StdRng::seed_from_u64(seed ^ (1ull << 63));
    StdRng fpu_rng = StdRng::seed_from_u64((seed ^ (1ull << 63)) + 1);
    RolloutPolicy<Connect4> rp(&rng, &cnt);
    Connect4Net net(weights, opt.libm);
    Connect4Net net2(weights2 ? weights2 : weights, opt.libm);
    CallbackPolicy cb(callback, ctx);
    Connect4 game;
    int first_player = game.player();
    uint32_t ply = 0;
    if (child_visits63x9) std::memset(child_visits63x9, 0, 63 * 36);
    for (;;) {
        const syn_player_cfg& pl = players[ply & 1u];
        uint32_t E = explores2 ? explores2[ply & 1u] : pl.num_explores;
        if (mover_out) *mover_out = ply & 1u;
        Policy<Connect4>* p = pl.leaf_eval_kind == SYN_LEAF_ROLLOUT ? (Policy<Connect4>*)&rp
                              : (callback ? (Policy<Connect4>*)&cb : ((ply & 1u) ? (Policy<Connect4>*)&net2 : (Policy<Connect4>*)&net));
        if (pl.leaf_eval_kind == SYN_LEAF_NN && !weights && !callback) return SYN_ERR_NO_WEIGHTS;
        int action;
        if (pl.tree_kind == SYN_TREE_MCTS) {
            MCTS<Connect4> m(E + 1, pl.mcts, p, game, opt, &noise_rng, &fpu_rng, &cnt);
            m.explore_n(E);
            m.finish();
            action = m.best_action(pl.action_selection);
            const auto& r = m.nodes[m.root];
            if (tree_nodes63) tree_nodes63[ply] = (uint32_t)m.nodes.size();
            if (child_visits63x9)
                for (uint32_t c = r.first_child; c < r.last_child(); ++c) child_visits63x9[9 * ply + m.nodes[c].action] = m.nodes[c].num_visits;
        } else {
            FrozenMCTS<Connect4> m(E + 1, pl.mcts, p, game, opt, &cnt);
            m.explore_n(E);
            m.finish();
            if (m.unsupported) return SYN_ERR_UNSUPPORTED;
            action = m.best_action(pl.action_selection);
            if (action < 0) return SYN_ERR_DEVICE_FAULT; // the reference's unwrap() panics
            const auto& r = m.nodes[m.root];
            if (tree_nodes63) tree_nodes63[ply] = (uint32_t)m.nodes.size();
            if (child_visits63x9)
                for (uint32_t c = r.first_child; c < r.last_child(); ++c) child_visits63x9[9 * ply + m.nodes[c].action] = m.nodes[c].num_visits;
        }
        if (moves63) moves63[ply] = (uint8_t)action;
        ++ply;
        if (game.step(action)) break;
    }
    if (result) *result = game.reward(first_player);
    if (n_moves) *n_moves = (uint8_t)ply;
    fill_stats(stats, cnt, ply, 1, 0);
    return 0;
}

int orc_match(const syn_player_cfg* players, uint64_t seed, const uint32_t* explores2, const float* weights, orc_eval_fn callback,
              void* ctx, uint32_t flags, float* result, uint8_t* n_moves, uint8_t* moves63, uint32_t* tree_nodes63,
              float* child_visits63x9, syn_stats* stats) {
    return orc_match2(players, seed, explores2, weights, nullptr, nullptr, callback, ctx, flags, result, n_moves, moves63, tree_nodes63,
                      child_visits63x9, stats);
}

// ---- gather with the engine's per-game streams.  trace_* (optional, one entry per row):
// the action played, nodes.len() of that ply's tree, and the root's child visit counts.
int orc_gather(const syn_rollout_cfg* cfg, const float* weights, orc_eval_fn callback, void* ctx, uint64_t seed,
               uint64_t first_game, uint32_t num_games, int threads, uint32_t flags, syn_experience* out, syn_stats* stats,
               uint8_t* trace_action, uint32_t* trace_nodes, float* trace_child_visits) {
    LeafSource leaf;
    leaf.weights = weights;
    leaf.callback = callback;
    leaf.callback_ctx = ctx;
    leaf.use_cache = (flags & ORC_FLAG_NO_CACHE) == 0;
    if (cfg->leaf_eval_kind == SYN_LEAF_NN && !weights && !callback) return SYN_ERR_NO_WEIGHTS;
    std::vector<GameTrace> traces;
    bool want_trace = trace_action || trace_nodes || trace_child_visits;
    GatherResult r = gather_streams(*cfg, leaf, seed, first_game, num_games, threads, opts_from(flags), want_trace ? &traces : nullptr);
    if (!r.ok) return SYN_ERR_INVALID_ARGUMENT;
    int rc = copy_experience(r.buffer, out);
    fill_stats(stats, r.counters, r.buffer.curr_steps(), num_games, r.elapsed_ns);
    if (rc) return rc;
    if (want_trace) {
        size_t row = 0;
        for (auto& t : traces)
            for (size_t k = 0; k < t.actions.size(); ++k, ++row) {
                if (trace_action) trace_action[row] = t.actions[k];
                if (trace_nodes) trace_nodes[row] = t.tree_nodes[k];
                if (trace_child_visits) std::memcpy(trace_child_visits + 9 * row, t.child_visits[k].data(), 36);
            }
    }
    return 0;
}

// ---- gather exactly as the reference schedules it (alpha_zero.rs:120-209); NN leaves only.
int orc_gather_reference(const syn_rollout_cfg* cfg, const float* weights, uint32_t num_workers, uint32_t games_per_train,
                         uint64_t seed, uint32_t flags, syn_experience* out, syn_stats* stats, uint64_t* cache_hits,
                         uint64_t* cache_misses) {
    if (!weights) return SYN_ERR_NO_WEIGHTS;
    if (cfg->leaf_eval_kind != SYN_LEAF_NN) return SYN_ERR_UNSUPPORTED;
    LeafSource leaf;
    leaf.weights = weights;
    leaf.use_cache = (flags & ORC_FLAG_NO_CACHE) == 0;
    GatherResult r = gather_reference_schedule(*cfg, leaf, num_workers, games_per_train, (size_t)seed, opts_from(flags));
    if (!r.ok) return SYN_ERR_INVALID_ARGUMENT;
    fill_stats(stats, r.counters, r.buffer.curr_steps(), games_per_train, r.elapsed_ns);
    if (cache_hits) *cache_hits = r.cache_hits;
    if (cache_misses) *cache_misses = r.cache_misses;
    if (out) return copy_experience(r.buffer, out);
    return 0;
}

// ---- ReplayBuffer::keep_last_n_games + extend on game-id arrays (data.rs:160-194): returns how many
// leading rows keep_last_n_games(n) removes from a buffer whose rows carry `game_ids` and whose
// game counter is `game_id`.
size_t orc_keep_last_n_games_prefix(const uint64_t* game_ids, size_t len, uint64_t game_id, uint64_t n) {
    if (game_id <= n) return 0;
    uint64_t min_game_id = game_id - n;
    size_t remove = 0;
    for (size_t i = 0; i < len; ++i) {
        if (game_ids[i] >= min_game_id) break;
        remove = i + 1;
    }
    return remove;
}

// ---- ReplayBuffer::deduplicate (data.rs:196-235): HashMap<G, StateStatistics> keyed by the position; sums run in
// buffer order (data.rs:199-215), averages are sum / num as f32 (data.rs:220-229).  The reference iterates the HashMap
// (random order per process); the groups are emitted here in order of first occurrence, which is what the device emits.
// Returns the number of distinct positions (which may exceed `capacity`: then nothing past capacity is written).
size_t orc_deduplicate(const uint64_t* my_bb, const uint64_t* op_bb, const float* pis, const float* vs, size_t n, size_t capacity,
                       float* out_states, float* out_pis, float* out_vs, uint64_t* out_my, uint64_t* out_op, uint32_t* out_num) {
    struct Stat { float sum_pi[9]; float sum_v[3]; uint32_t num; size_t first; };
    struct KeyHash { size_t operator()(const std::pair<uint64_t, uint64_t>& k) const { return (size_t)(k.first * 0x9E3779B97F4A7C15ull ^ (k.second + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull); } };
    std::unordered_map<std::pair<uint64_t, uint64_t>, size_t, KeyHash> index;
    std::vector<Stat> stats;
    index.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        auto key = std::make_pair(my_bb[i], op_bb[i]);
        auto it = index.find(key);
        if (it == index.end()) {
            it = index.emplace(key, stats.size()).first;
            Stat st;
            std::memset(&st, 0, sizeof(st));
            st.first = i;
            stats.push_back(st);
        }
        Stat& st = stats[it->second];
        for (int j = 0; j < 9; ++j) st.sum_pi[j] += pis[9 * i + j];
        for (int j = 0; j < 3; ++j) st.sum_v[j] += vs[3 * i + j];
        st.num += 1;
    }
    for (size_t g = 0; g < stats.size() && g < capacity; ++g) {
        const Stat& st = stats[g];
        for (int j = 0; j < 9; ++j) if (out_pis) out_pis[9 * g + j] = st.sum_pi[j] / (float)st.num;
        for (int j = 0; j < 3; ++j) if (out_vs) out_vs[3 * g + j] = st.sum_v[j] / (float)st.num;
        if (out_my) out_my[g] = my_bb[st.first];
        if (out_op) out_op[g] = op_bb[st.first];
        if (out_num) out_num[g] = st.num;
        if (out_states) {
            Connect4::from_bitboards(my_bb[st.first], op_bb[st.first]).features(out_states + 63 * g);
        }
    }
    return stats.size();
}

} // extern "C"
