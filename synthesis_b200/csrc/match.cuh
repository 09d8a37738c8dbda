// match.cuh — evaluation matches and the evaluator's FrozenMCTS, thread per match.
//
// Replaces synthesis/src/evaluator.rs:129-228 (eval_against_old, eval_against_rollout_mcts,
// mcts_vs_mcts) and :299-534 (FrozenMCTS) of the reference.  A thread plays a whole match: every
// move it builds a fresh tree of the mover's kind (tpg.cuh's MCTS or the FrozenMCTS below),
// explores it and plays best_action.  Leaves of a Connect4Net player go through the team's
// tensor-core forward exactly like self-play (mlp_team.cuh); leaves of a RolloutPolicy player are
// rolled out by the thread itself on the match's ChaCha12 stream.  One explore per round, so the
// team's barrier is never held for longer than one descent + one rollout.
//
// FrozenMCTS reuses the 32-byte node record of tpg.cuh with word 1 = cum_value (words 2, 3 unused).
#pragma once
#include "mlp_split.cuh"
#include "mlp_team.cuh"
#include "tpg.cuh"

namespace frz {

using namespace tpg;

// Outcome::value (game.rs:37-43)
__device__ __forceinline__ float sol_value(uint32_t s) {
    uint32_t k = sol_kind(s);
    return k == SYN_KIND_WIN ? 1.0f : (k == SYN_KIND_LOSE ? -1.0f : 0.0f);
}

// evaluator.rs:482-527
__device__ __forceinline__ void backprop(const syn_mcts_cfg& cfg, Game& g, uint32_t id, float value, bool solved) {
    uint4* nodes = g.nodes;
    uint32_t levels = 0;
    for (;;) {
        Rec n = load_rec(nodes, id);
        ++levels;
        uint32_t nch = n.pk & 0xffu, nsol = (n.pk >> 8) & 0xffu;
        if (cfg.solve && solved && nsol == 0u) {
            bool all_solved = true;
            uint32_t worst = 0u;
            for (uint32_t k = 0; k < nch; ++k) {
                uint32_t csol = (meta_words(nodes, n.fc + k)[3] >> 8) & 0xffu;
                if (csol == 0u) all_solved = false; // unvisited or unsolved
                else if (worst == 0u || sol_key(csol) < sol_key(worst)) worst = csol;
            }
            if (sol_kind(worst) == SYN_KIND_LOSE) { // a child is lost for its mover: this node is a win
                meta_words(nodes, id)[3] = (n.pk & 0xffff00ffu) | (c4::SOL_WIN0 << 8);
                value = __fadd_rn(-n.o0, __fadd_rn(n.vis, 1.0f));
            } else if (nch != 0u && all_solved) {
                uint32_t best = sol_reversed(worst);
                meta_words(nodes, id)[3] = (n.pk & 0xffff00ffu) | (best << 8);
                value = sol_kind(best) == SYN_KIND_DRAW ? -n.o0 : __fsub_rn(-n.o0, __fadd_rn(n.vis, 1.0f));
            } else {
                solved = false;
            }
        }
        store_stat(nodes, id, __fadd_rn(n.vis, 1.0f), __fadd_rn(n.o0, value), 0.0f, 0.0f);
        value = -value;
        if (id == 0u) break;
        id = n.parent;
    }
    g.cnt[CNT_BACKPROP_LEVELS] += levels;
}

// One explore up to the point where the policy is needed (evaluator.rs:382-438 and the child loop
// of visit, :446-463).  Returns true with `lf` filled when a leaf awaits Policy::eval.  Only
// Fpu::Const and Exploration::Uct exist here (the reference panics otherwise, :410, :424; the
// host rejects such configs).
__device__ __forceinline__ bool descend(const syn_mcts_cfg& cfg, uint32_t cap, Game& g, uint64_t& my, uint64_t& op, Leaf& lf) {
    uint4* nodes = g.nodes;
    uint32_t cur = 0u, depth = 0u;
    float cvis;
    uint32_t cfc, cpk;
    {
        Rec r = load_rec(nodes, 0u);
        cvis = r.vis; cfc = r.fc; cpk = r.pk;
    }
    for (;;) {
        uint32_t sol = (cpk >> 8) & 0xffu, nch = cpk & 0xffu;
        if (sol) { backprop(cfg, g, cur, sol_value(sol), true); return false; } // :386-388
        if (nch == 0u) break;
        const float pterm = __fsqrt_rn(__fmul_rn(cfg.c, syn_logf(cvis)));
        uint32_t b = 0u, bfc = 0u, bpk = 0u;
        float bval = 0.0f, bvis = 0.0f;
        // five children per trip: their records are requested together (a level costs ceil(nch / 5) memory round trips
        // instead of nch), scored in child order like the reference's loop
        for (uint32_t k0 = 0; k0 < nch; k0 += 5u) {
            Rec chs[5];
#pragma unroll
            for (uint32_t j = 0; j < 5u; ++j) chs[j] = load_rec(nodes, cfc + (k0 + j < nch ? k0 + j : nch - 1u));
#pragma unroll
            for (uint32_t j = 0; j < 5u; ++j) {
                const uint32_t k = k0 + j;
                if (k < nch) {
                    const Rec& ch = chs[j];
                    uint32_t csol = (ch.pk >> 8) & 0xffu, cn = ch.pk & 0xffu;
                    float value;
                    if (csol == 0u && cn == 0u) {
                        value = __fadd_rn(cfg.fpu_a, __uint_as_float(ch.prior));
                    } else {
                        float q = csol ? sol_value(sol_reversed(csol)) : -__fdiv_rn(ch.o0, ch.vis);
                        value = __fadd_rn(q, __fdiv_rn(pterm, __fsqrt_rn(ch.vis)));
                    }
                    if (k == 0u || value > bval) { b = k; bval = value; bvis = ch.vis; bfc = ch.fc; bpk = ch.pk; }
                }
            }
        }
        g.cnt[CNT_SELECT_LEVELS] += 1u;
        g.cnt[CNT_CHILDREN_SCANNED] += nch;
        cur = cfc + b;
        cvis = bvis; cfc = bfc; cpk = bpk;
        c4::step(my, op, (int)((cpk >> 16) & 0xffu));
        if (++depth >= 64u) { g.err = DERR_DEPTH_OVERFLOW; return false; }
    }
    uint64_t occ = my | op;
    uint32_t lm = 0u, cs2 = 0u, n = 0u;
#pragma unroll
    for (int col = 0; col < 9; ++col) {
        uint32_t colbits = (uint32_t)((occ >> (7 * col)) & 0x7full);
        if (colbits != 0x7fu) {
            uint64_t bit = 1ull << (7 * col + __popc(colbits));
            uint32_t s2 = c4::won(my | bit) ? 1u : (((occ | bit) == c4::ALL) ? 2u : 0u);
            lm |= 1u << col;
            cs2 |= s2 << (2 * col);
            ++n;
        }
    }
    uint32_t fc = g.nn;
    if (fc + n > cap) { g.err = DERR_ARENA_OVERFLOW; return false; }
    uint32_t* mw = meta_words(nodes, cur);
    mw[1] = fc;
    mw[3] = (cpk & 0xffffff00u) | n;
    g.nn = fc + n;
    g.cnt[CNT_EXPANSIONS] += 1u;
    g.cnt[CNT_CHILDREN_CREATED] += n;
    lf.id = cur; lf.fc = fc; lf.legal = lm; lf.csol2 = cs2; lf.any_solved = cs2 != 0u;
    return true;
}

// The rest of visit() (evaluator.rs:446-479): child records with softmax priors over the legal
// children, value = dist[2] - dist[0], then backprop.
__device__ __forceinline__ void finish(const syn_mcts_cfg& cfg, Game& g, const Leaf& lf, bool uniform_prior, const float (&logits)[9],
                                       float o0, float o2) {
    uint4* nodes = g.nodes;
    const uint32_t n = (uint32_t)__popc(lf.legal);
    float e[9];
    float total = 0.0f;
    if (!uniform_prior) {
        float mx = __uint_as_float(0xff800000u);
#pragma unroll
        for (int col = 0; col < 9; ++col)
            if ((lf.legal >> col) & 1u) mx = fmaxf(mx, logits[col]);
#pragma unroll
        for (int col = 0; col < 9; ++col) {
            e[col] = 0.0f;
            if ((lf.legal >> col) & 1u) {
                e[col] = syn_expf(__fsub_rn(logits[col], mx));
                total = __fadd_rn(total, e[col]);
            }
        }
    }
    const float up = __fdiv_rn(1.0f, (float)n);
    uint32_t rank = 0u;
#pragma unroll
    for (int col = 0; col < 9; ++col) {
        if ((lf.legal >> col) & 1u) {
            uint32_t s2 = (lf.csol2 >> (2 * col)) & 3u;
            uint32_t csol = s2 == 1u ? c4::SOL_LOSE0 : (s2 == 2u ? c4::SOL_DRAW0 : 0u);
            float prior = uniform_prior ? up : __fdiv_rn(e[col], total);
            store_new_child(nodes, lf.fc + rank, prior, lf.id, (csol << 8) | ((uint32_t)col << 16));
            ++rank;
        }
    }
    g.cnt[CNT_LEAF_EVALS] += 1u;
    backprop(cfg, g, lf.id, __fsub_rn(o2, o0), lf.any_solved);
}

// evaluator.rs:356-380 plus what the parity tests read.  best_action = -1: every child unvisited
// (the reference's unwrap() panics).
__device__ __noinline__ void read_root(const Game& g, uint32_t action_selection, RootOut& r) {
    const uint4* nodes = g.nodes;
    Rec root = load_rec(nodes, 0u);
    uint32_t nch = root.pk & 0xffu;
    for (int k = 0; k < 9; ++k) { r.pi[k] = 0.0f; r.visits[k] = 0.0f; r.child_sol[k] = 0u; }
    int best = -1;
    float bv = __uint_as_float(0xff800000u);
    uint32_t legal = 0u;
    for (uint32_t k = 0; k < nch; ++k) {
        Rec ch = load_rec(nodes, root.fc + k);
        uint32_t csol = (ch.pk >> 8) & 0xffu, cn = ch.pk & 0xffu, act = (ch.pk >> 16) & 0xffu;
        legal |= 1u << act;
#pragma unroll
        for (int col = 0; col < 9; ++col)
            if ((int)act == col) { r.visits[col] = ch.vis; r.child_sol[col] = csol; }
        if (csol == 0u && cn == 0u) continue;
        uint32_t kind = sol_kind(csol);
        float value;
        if (kind == SYN_KIND_WIN) value = __uint_as_float(0xff800000u);
        else if (kind == SYN_KIND_DRAW) value = 1e6f;
        else if (kind == SYN_KIND_LOSE) value = __uint_as_float(0x7f800000u);
        else value = action_selection == SYN_ACTION_Q ? __fdiv_rn(-ch.o0, ch.vis) : ch.vis;
        if (best < 0 || value > bv) { bv = value; best = (int)act; }
    }
    r.legal = legal;
    r.best_action = best;
    r.root_sol = (root.pk >> 8) & 0xffu;
    r.q0 = root.o0; r.q1 = 0.0f; r.q2 = root.vis; // FrozenMCTS keeps a scalar: [cum_value, 0, num_visits]
}

} // namespace frz

namespace mtc {

using namespace tpg;

// policies/rollout.rs:8-31 by one thread on the match's stream.  Returns the one-hot index
// (0 Lose, 1 Draw, 2 Win) for the leaf's player to move.
__device__ __noinline__ int rollout(rng::Stream& st, uint64_t my, uint64_t op, uint32_t& plies) {
    uint32_t k = 0;
    for (;;) {
        uint64_t occ = my | op;
        uint64_t legal = (~(occ >> 6)) & c4::ROW0; // bit 7c set <=> column c has room
        uint32_t n = (uint32_t)__popcll(legal);
        uint32_t hi = st.gen_range_1to9(n);
        for (uint32_t t = 0; t < hi; ++t) legal &= legal - 1; // hi-th legal column, ascending
        int p7 = __ffsll((long long)legal) - 1;
        uint64_t bit = (occ + (1ull << p7)) & (0x7full << p7);
        uint64_t mover = my | bit;
        my = op;
        op = mover;
        ++k;
        if (c4::won(mover)) { plies += k; return (k & 1u) ? 2 : 0; }
        if ((occ | bit) == c4::ALL) { plies += k; return 1; }
    }
}

struct MParams { // what a match launch adds to KParams
    syn_player_cfg players[2];
    const uint32_t* explores; // [n][2] or null
    float* result;            // [n]
    uint8_t* n_moves;         // [n]
    uint32_t active_per_block; // threads of a CTA that play (arena slot = block * active + thread)
    const uint8_t* weight_image2; // players[1]'s network when it differs from players[0]'s (eval_against_old), else null
};

__device__ __forceinline__ const syn_player_cfg& mover(const MParams& m, const Game& g) { return m.players[g.ply & 1u]; }
__device__ __forceinline__ uint32_t mover_explores(const KParams& p, const MParams& m, const Game& g) {
    if (m.explores && !p.search_mode) return m.explores[2 * (size_t)g.gi + (g.ply & 1u)];
    return mover(m, g).num_explores;
}

// Ends the current move (MCTS::exploit / FrozenMCTS::exploit return best_action; the game loops of
// evaluator.rs:173-197 / 211-227 step the game and stop when it is over).  Cold path.
__device__ __noinline__ void end_of_move(const KParams& p, const MParams& m, Game& g) {
    const syn_player_cfg& pl = mover(m, g);
    RootOut r;
    if (pl.tree_kind == SYN_TREE_FROZEN) frz::read_root(g, pl.action_selection, r);
    else tpg::read_root(g, pl.action_selection, r);
    g.cnt[CNT_NODES] += g.nn;
    if (p.search_mode) {
        size_t i = g.gi;
        for (int k = 0; k < 9; ++k) {
            if (p.s_child_visits) p.s_child_visits[i * 9 + k] = r.visits[k];
            if (p.s_child_sol) p.s_child_sol[i * 9 + k] = (uint8_t)r.child_sol[k];
        }
        if (p.s_root_q) { p.s_root_q[i * 3 + 0] = r.q0; p.s_root_q[i * 3 + 1] = r.q1; p.s_root_q[i * 3 + 2] = r.q2; }
        if (p.s_root_sol) p.s_root_sol[i] = (uint8_t)r.root_sol;
        if (p.s_best) p.s_best[i] = (uint8_t)(r.best_action < 0 ? 255 : r.best_action);
        if (p.s_nodes) p.s_nodes[i] = g.nn;
        g.cnt[CNT_GAMES] += 1u;
        g.phase = PH_NEED_GAME;
        return;
    }
    if (r.best_action < 0) { g.err = DERR_NO_BEST_ACTION; return; }
    size_t row = (size_t)g.gi * 63 + g.ply;
    for (int k = 0; k < 9; ++k) p.row_visits[row * 9 + k] = r.visits[k];
    p.row_action[row] = (uint8_t)r.best_action;
    p.row_nodes[row] = g.nn;
    uint32_t over = c4::step(g.my, g.op, r.best_action);
    g.ply += 1;
    if (over == 0u) { g.phase = PH_NEW_TREE; return; }
    // game.reward(first_player): the last mover (ply - 1) won unless the board filled up
    float res = 0.0f;
    if (over == c4::SOL_LOSE0) res = ((g.ply - 1u) & 1u) ? -1.0f : 1.0f;
    m.result[g.gi] = res;
    m.n_moves[g.gi] = (uint8_t)g.ply;
    g.cnt[CNT_ROWS] += g.ply;
    g.cnt[CNT_GAMES] += 1u;
    g.phase = PH_NEED_GAME;
}

// Runs the thread's state machine for at most ONE explore.  Returns 0 when no matches are left
// (phase == PH_DONE), 1 when a leaf awaits the network (lf filled, board in my/op), 2 otherwise.
// Matches are dealt statically, match i to thread (i / gridDim) of CTA (i % gridDim), so that a small
// batch spreads over all SMs.
__device__ __forceinline__ int advance(const KParams& p, const MParams& m, Game& g, uint32_t& next_gi, rng::Stream& rs, Leaf& lf, uint64_t& my,
                                       uint64_t& op) {
    if (g.phase == PH_DONE) return 0;
    if (g.phase == PH_NEED_GAME) {
        flush_counters(p, g);
        uint32_t gi = next_gi;
        next_gi += gridDim.x * m.active_per_block;
        if (gi >= p.num_games || *(volatile int*)p.error != 0) { g.phase = PH_DONE; return 0; }
        g.gi = gi;
        g.ply = 0; g.apos = 0; g.fpu_pos = 0; g.noise_pos = 0;
        if (p.search_mode) { g.my = p.pos_my[gi]; g.op = p.pos_op[gi]; }
        else { g.my = 0; g.op = 0; }
        g.stream = p.pos_seed[gi];
        rs.init(g.stream, 0);
        g.phase = PH_NEW_TREE;
    }
    const syn_player_cfg& pl = mover(m, g);
    const bool frozen = pl.tree_kind == SYN_TREE_FROZEN;
    if (g.phase == PH_NEW_TREE) { // with_capacity (mcts.rs:123-137, evaluator.rs:320-333): fresh arena, root only
        g.nodes[0] = make_uint4(0u, 0u, 0u, 0u);
        g.nodes[1] = make_uint4(0u, 0u, 0u, 0u);
        g.nn = 1u;
        g.e_done = 0u;
        g.is_init = true;
        g.cnt[CNT_TREES] += 1u;
        g.phase = PH_EXPLORE;
    } else { // explore_n: MCTS stops when the root is solved (mcts.rs:139-147), FrozenMCTS does not (evaluator.rs:529-533)
        uint32_t rpk = meta_words(g.nodes, 0u)[3];
        if (g.e_done >= mover_explores(p, m, g) || (!frozen && ((rpk >> 8) & 0xffu) != 0u)) {
            end_of_move(p, m, g);
            if (g.err) { fail(p, g); return 0; }
            return 2;
        }
    }
    if (!g.is_init) g.cnt[CNT_EXPLORES] += 1u;
    Ctx c;
    c.cfg = &pl.mcts; c.cap = p.arena_nodes; c.seed = 0; c.search_mode = true; // streams: rollout = seed, noise/fpu derived (tpg::Ctx)
    my = g.my; op = g.op;
    bool need = frozen ? frz::descend(pl.mcts, p.arena_nodes, g, my, op, lf) : tpg::descend(c, g, my, op, lf);
    if (g.err) { fail(p, g); return 0; }
    if (need && pl.leaf_eval_kind == SYN_LEAF_NN) return 1;
    if (need) { // RolloutPolicy: logits = 0, one-hot outcome (policies/rollout.rs:24-30)
        uint32_t plies = 0;
        int idx = rollout(rs, my, op, plies);
        g.cnt[CNT_ROLLOUT_PLIES] += plies;
        const float zero[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float o0 = idx == 0 ? 1.0f : 0.0f, o1 = idx == 1 ? 1.0f : 0.0f, o2 = idx == 2 ? 1.0f : 0.0f;
        if (frozen) frz::finish(pl.mcts, g, lf, true, zero, o0, o2);
        else tpg::finish(c, g, lf, true, zero, o0, o1, o2);
    }
    if (g.is_init) { if (!frozen) add_root_noise(c, g); g.is_init = false; }
    else g.e_done += 1u;
    return 2;
}

// After the network evaluated the leaf returned by advance() == 1.
__device__ __forceinline__ void finish_nn(const KParams& p, const MParams& m, Game& g, const Leaf& lf, const float (&y)[12]) {
    const syn_player_cfg& pl = mover(m, g);
    float mx = fmaxf(y[9], fmaxf(y[10], y[11])); // value.softmax(-1) (study-connect4/src/policies.rs:54-56)
    float e0 = syn_expf(__fsub_rn(y[9], mx)), e1 = syn_expf(__fsub_rn(y[10], mx)), e2 = syn_expf(__fsub_rn(y[11], mx));
    float tot = __fadd_rn(__fadd_rn(e0, e1), e2);
    float lg[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) lg[k] = y[k];
    Ctx c;
    c.cfg = &pl.mcts; c.cap = p.arena_nodes; c.seed = 0; c.search_mode = true;
    if (pl.tree_kind == SYN_TREE_FROZEN) {
        frz::finish(pl.mcts, g, lf, false, lg, __fdiv_rn(e0, tot), __fdiv_rn(e2, tot));
        if (g.is_init) g.is_init = false;
        else g.e_done += 1u;
    } else {
        tpg::finish(c, g, lf, false, lg, __fdiv_rn(e0, tot), __fdiv_rn(e1, tot), __fdiv_rn(e2, tot));
        tpg::after_eval(c, g);
    }
}

} // namespace mtc

namespace eng {

// One persistent CTA per SM, TEAMS teams of 128 threads; thread = match (or search root).
// TWO: players[1] evaluates its leaves with a second network (eval_against_old, evaluator.rs:129-160, with p1 != p2):
// both images stay resident, and a round in which a team holds leaves of both players runs the forward once per image.
template <int TEAMS, int SLOTS, bool TWO>
__global__ void __launch_bounds__(128 * TEAMS, 1) match_tpg_kernel(const __grid_constant__ KParams p, const __grid_constant__ mtc::MParams m) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    mlpteam::Smem<TEAMS, SLOTS>& ms = *reinterpret_cast<mlpteam::Smem<TEAMS, SLOTS>*>(smem_raw);
    uint8_t* const img2 = smem_raw + sizeof(mlpteam::Smem<TEAMS, SLOTS>);
    const bool any_nn = m.players[0].leaf_eval_kind == SYN_LEAF_NN || m.players[1].leaf_eval_kind == SYN_LEAF_NN;
    if (any_nn) mlpteam::setup<TEAMS, SLOTS>(ms, p.weight_image);
    if (TWO) mlpteam::load_second_image<TEAMS, SLOTS>(ms, img2, m.weight_image2);
    const int team = threadIdx.x >> 7, r = threadIdx.x & 127;
    tpg::Game g;
    tpg::init_game(p, g, (size_t)blockIdx.x * m.active_per_block + (threadIdx.x < m.active_per_block ? threadIdx.x : 0u));
    if (threadIdx.x >= m.active_per_block) g.phase = PH_DONE;
    uint32_t next_gi = threadIdx.x * gridDim.x + blockIdx.x;
    rng::Stream rs;
    rs.init(0, 0);
    tpg::Leaf lf;
    uint64_t my = 0, op = 0;
    for (;;) {
        int st = mtc::advance(p, m, g, next_gi, rs, lf, my, op);
        __syncwarp();
        if (!mlpteam::team_any(team, st != 0)) break; // no thread of this team has a match left
        if (!any_nn) continue;
        const bool second = TWO && st == 1 && (g.ply & 1u) != 0u; // this leaf belongs to players[1]'s tree
        const bool any1 = mlpteam::team_any(team, st == 1 && !second);
        const bool any2 = TWO && mlpteam::team_any(team, second);
        if (!any1 && !any2) continue; // nobody needs a network this round
        uint32_t mma_phase;
        const int slot = mlpteam::acquire_slot<TEAMS, SLOTS>(ms, team, r, mma_phase);
        float y[12];
        if (any1) {
            if (st == 1) mlpteam::write_features(ms.a[slot], ms.col_lut, r, my, op);
            mlpteam::forward<TEAMS, SLOTS>(ms, team, slot, r, mma_phase, y);
        }
        if (any2) { // the chain overwrites its A tile in place, so the features are written again
            float y2[12];
            if (st == 1) mlpteam::write_features(ms.a[slot], ms.col_lut, r, my, op);
            mlpteam::forward_img<TEAMS, SLOTS>(ms, img2, team, slot, r, mma_phase, y2);
            if (second) {
#pragma unroll
                for (int k = 0; k < 12; ++k) y[k] = y2[k];
            }
        }
        mlpteam::release_slot<TEAMS, SLOTS>(ms, team, r, slot, mma_phase);
        if (st == 1) mtc::finish_nn(p, m, g, lf, y);
        __syncwarp();
    }
    tpg::flush_counters(p, g);
    if (any_nn) mlpteam::teardown<TEAMS, SLOTS>(ms);
}

// The same with the fp32-grade forward of mlp_split.cuh (one network: two split images do not fit one SM's shared memory, so
// matches between two DIFFERENT networks run match_tpg_kernel<.., TWO> with the single-fp16 chain).
template <int TEAMS>
__global__ void __launch_bounds__(128 * TEAMS, 1) match_tpg_split_kernel(const __grid_constant__ KParams p, const __grid_constant__ mtc::MParams m) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int SLOTS = TEAMS < 2 ? TEAMS : 2;
    constexpr bool TWO = false;
    mlps::Smem<TEAMS, SLOTS>& ms = *reinterpret_cast<mlps::Smem<TEAMS, SLOTS>*>(smem_raw);
    const bool any_nn = m.players[0].leaf_eval_kind == SYN_LEAF_NN || m.players[1].leaf_eval_kind == SYN_LEAF_NN;
    if (any_nn) mlps::setup<TEAMS, SLOTS>(ms, p.weight_image, p.weight_image_lo);
    const int team = threadIdx.x >> 7, r = threadIdx.x & 127;
    tpg::Game g;
    tpg::init_game(p, g, (size_t)blockIdx.x * m.active_per_block + (threadIdx.x < m.active_per_block ? threadIdx.x : 0u));
    if (threadIdx.x >= m.active_per_block) g.phase = PH_DONE;
    uint32_t next_gi = threadIdx.x * gridDim.x + blockIdx.x;
    rng::Stream rs;
    rs.init(0, 0);
    tpg::Leaf lf;
    uint64_t my = 0, op = 0;
    for (;;) {
        int st = mtc::advance(p, m, g, next_gi, rs, lf, my, op);
        __syncwarp();
        if (!mlps::team_any(team, st != 0)) break; // no thread of this team has a match left
        if (!any_nn) continue;
        const bool second = TWO && st == 1 && (g.ply & 1u) != 0u; // this leaf belongs to players[1]'s tree
        const bool any1 = mlps::team_any(team, st == 1 && !second);
        const bool any2 = TWO && mlps::team_any(team, second);
        if (!any1 && !any2) continue; // nobody needs a network this round
        uint32_t mma_phase;
        const int slot = mlps::acquire_slot<TEAMS, SLOTS>(ms, team, r, mma_phase);
        float y[12];
        mlps::write_features<TEAMS, SLOTS>(ms, slot, r, my, op, st == 1);
        mlps::forward<TEAMS, SLOTS>(ms, p.mlp_bias, team, slot, r, mma_phase, y);
        mlps::release_slot<TEAMS, SLOTS>(ms, team, r, slot, mma_phase);
        if (st == 1) mtc::finish_nn(p, m, g, lf, y);
        __syncwarp();
    }
    tpg::flush_counters(p, g);
    if (any_nn) mlps::teardown<TEAMS, SLOTS>(ms);
}

} // namespace eng
