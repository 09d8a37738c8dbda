// tpg.cuh — thread-per-game MCTS, scalar device functions on whole 32-byte node records: the first
// thread-per-game implementation.  The self-play kernel built on it was replaced by tpg2.cuh (round-
// synchronous schedule); these functions remain the MCTS half of the evaluation-match kernel (match.cuh).
//
// Replaces synthesis/src/mcts.rs:29-489 and synthesis/src/alpha_zero.rs:229-338 of the
// reference, one THREAD per game.  The reference's loops over <= 9 children stay loops, in the
// reference's order, so every f32 operation happens in the same sequence as on the CPU; what is
// parallel is the 32 games of a warp and the 128 games of a team, whose leaves form one
// 128-row tile for the tensor-core Connect4Net forward (mlp_team.cuh) — thread r of a team owns
// tile row r from the feature write to the TMEM read-back, so a leaf's logits land in the
// registers of the thread that needs them.
//
// A tree is still a strictly serial object (one explore at a time, single writer): visit counts
// stay bit-identical to the reference.  Node records: tree.cuh (32 bytes, one sector).
#pragma once
#include "selfplay.cuh"

namespace tpg {

using namespace eng;

struct Rec { // one node record in registers
    float vis, o0, o1, o2;
    uint32_t prior, fc, parent, pk; // pk = num_children | solution << 8 | action << 16
};

__device__ __forceinline__ Rec load_rec(const uint4* nodes, uint32_t i) { // one 256-bit load (LDG.E.256) = one sector
    Rec r;
    unsigned long long q0, q1, q2, q3;
    asm volatile("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(nodes + 2 * (size_t)i) : "memory");
    r.vis = __uint_as_float((uint32_t)q0); r.o0 = __uint_as_float((uint32_t)(q0 >> 32));
    r.o1 = __uint_as_float((uint32_t)q1); r.o2 = __uint_as_float((uint32_t)(q1 >> 32));
    r.prior = (uint32_t)q2; r.fc = (uint32_t)(q2 >> 32); r.parent = (uint32_t)q3; r.pk = (uint32_t)(q3 >> 32);
    return r;
}
__device__ __forceinline__ void store_stat(uint4* nodes, uint32_t i, float vis, float o0, float o1, float o2) {
    nodes[2 * (size_t)i] = make_uint4(__float_as_uint(vis), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2));
}
__device__ __forceinline__ void store_new_child(uint4* nodes, uint32_t i, float prior, uint32_t parent, uint32_t pk) {
    nodes[2 * (size_t)i] = make_uint4(0u, 0u, 0u, 0u);
    nodes[2 * (size_t)i + 1] = make_uint4(__float_as_uint(prior), 0u, parent, pk);
}
__device__ __forceinline__ uint32_t* meta_words(uint4* nodes, uint32_t i) { return reinterpret_cast<uint32_t*>(nodes + 2 * (size_t)i + 1); }

// Per-thread game + tree state (registers).
struct Game {
    uint4* nodes;
    uint32_t nn;       // nodes.len()
    uint64_t my, op;   // root position of the current tree
    uint64_t stream;   // global game index (gather) or the position's rollout seed (search)
    uint32_t gi, ply, e_done, apos, fpu_pos, noise_pos;
    int phase;
    bool is_init;
    int err;
    uint32_t cnt[CNT_N];
};

struct Leaf { // an expanded leaf waiting for Policy::eval; its board is left in (my, op) of the caller
    uint32_t id, fc, legal, csol2; // csol2: 2 bits per column, 0 none / 1 Lose(0) / 2 Draw(0)
    bool any_solved;
};

struct Ctx {
    const syn_mcts_cfg* cfg;
    uint32_t cap;
    uint64_t seed;
    bool search_mode;
    __device__ __forceinline__ uint64_t stream_seed(const Game& g, unsigned k) const {
        if (!search_mode) return syn_stream_seed(seed, g.stream, k);
        if (k == SYN_STREAM_ROLLOUT) return g.stream;
        if (k == SYN_STREAM_ACTION) return 0ull;
        return (g.stream ^ (1ull << 63)) + (k == SYN_STREAM_FPU ? 1ull : 0ull);
    }
};

// mcts.rs:429-488 from node `id` up to the root, following parent links like the reference.
__device__ __forceinline__ void backprop(const Ctx& c, Game& g, uint32_t id, float v0, float v1, float v2, bool solved) {
    const syn_mcts_cfg& cfg = *c.cfg;
    uint4* nodes = g.nodes;
    uint32_t levels = 0;
    for (;;) {
        Rec n = load_rec(nodes, id);
        ++levels;
        if (cfg.solve && solved) {
            uint32_t nch = n.pk & 0xffu, nsol = (n.pk >> 8) & 0xffu;
            uint32_t bk = sol_key(nsol);
            bool all_solved = true;
            for (uint32_t k = 0; k < nch; ++k) {
                uint32_t csol = (meta_words(nodes, n.fc + k)[3] >> 8) & 0xffu;
                uint32_t rs = csol ? sol_reversed(csol) : 0u;
                all_solved = all_solved && rs != 0u;
                uint32_t key = sol_key(rs);
                bk = key > bk ? key : bk;
            }
            uint32_t best = sol_from_key(bk);
            bool mark = false;
            int slot = 0;
            if (sol_kind(best) == SYN_KIND_WIN) { mark = true; slot = 2; }
            else if (best != 0u && all_solved) { mark = true; slot = sol_kind(best) == SYN_KIND_DRAW ? 1 : 0; }
            if (mark) {
                if (cfg.correct_values_on_solve) {
                    v0 = -n.o0; v1 = -n.o1; v2 = -n.o2;
                    float add = n.vis + 1.0f;
                    if (slot == 2) v2 = v2 + add;
                    else if (slot == 1) v1 = v1 + add;
                    else v0 = v0 + add;
                }
                meta_words(nodes, id)[3] = (n.pk & 0xffff00ffu) | (best << 8);
            } else {
                solved = false;
            }
        }
        store_stat(nodes, id, n.vis + 1.0f, n.o0 + v0, n.o1 + v1, n.o2 + v2);
        if (id == 0u) break;
        float tmp = v0; v0 = v2; v2 = tmp;
        id = n.parent;
    }
    g.cnt[CNT_BACKPROP_LEVELS] += levels;
}

__device__ __forceinline__ void backprop_onehot(const Ctx& c, Game& g, uint32_t id, uint32_t sol) {
    int idx = sol_index(sol);
    backprop(c, g, id, idx == 0 ? 1.0f : 0.0f, idx == 1 ? 1.0f : 0.0f, idx == 2 ? 1.0f : 0.0f, true);
}

__device__ __noinline__ float fpu_normal_draw(const Ctx& c, Game& g) { // mcts.rs:354 with the shipped closure; cold
    rng::Stream st;
    st.init(c.stream_seed(g, SYN_STREAM_FPU), g.fpu_pos);
    float v = syn_normal(st, c.cfg->fpu_a, c.cfg->fpu_b);
    g.fpu_pos = (uint32_t)st.pos;
    return v;
}

// One explore up to the point where the policy is needed (mcts.rs:310-325, 327-372, 374-406).
// `my`/`op` enter as the root position and leave as the leaf's position.  Returns true and fills
// `lf` if a leaf awaits evaluation (its children are numbered but NOT yet written: finish()
// writes each child record once, prior included); false if the explore is complete.
__device__ __forceinline__ bool descend(const Ctx& c, Game& g, uint64_t& my, uint64_t& op, Leaf& lf) {
    const syn_mcts_cfg& cfg = *c.cfg;
    uint4* nodes = g.nodes;
    uint32_t cur = 0u;
    float cvis, cop0, cop2;
    uint32_t cfc, cpk;
    {
        Rec r = load_rec(nodes, 0u);
        cvis = r.vis; cop0 = r.o0; cop2 = r.o2; cfc = r.fc; cpk = r.pk;
    }
    uint32_t depth = 0;
    for (;;) {
        uint32_t sol = (cpk >> 8) & 0xffu, nch = cpk & 0xffu;
        if (sol) { backprop_onehot(c, g, cur, sol); return false; } // mcts.rs:314-316
        if (nch == 0u) break;
        // ---- select_best_child (mcts.rs:327-372): first strict maximum in child order
        const bool puct = cfg.exploration_kind == SYN_EXPLORATION_POLYNOMIAL_UCT;
        const float pterm = puct ? __fsqrt_rn(cvis) : __fsqrt_rn(__fmul_rn(cfg.c, syn_logf(cvis)));
        uint32_t b = 0u, bfc = 0u, bpk = 0u;
        float bval = 0.0f, bvis = 0.0f, bo0 = 0.0f, bo2 = 0.0f;
        // five children per trip: their records are requested together, so a level costs
        // ceil(nch / 5) memory round trips instead of nch (the match kernel runs at 128 registers per thread)
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
                    float q;
                    if (csol) {
                        uint32_t kd = sol_kind(csol);
                        q = cfg.select_solved_nodes ? (kd == SYN_KIND_WIN ? -1.0f : (kd == SYN_KIND_LOSE ? 1.0f : 0.0f)) : __uint_as_float(0xff800000u);
                    } else if (cn == 0u) {
                        q = cfg.fpu_kind == SYN_FPU_CONST ? cfg.fpu_a : (cfg.fpu_kind == SYN_FPU_PARENT_Q ? __fdiv_rn(__fsub_rn(cop2, cop0), cvis) : fpu_normal_draw(c, g));
                    } else {
                        q = -__fdiv_rn(__fsub_rn(ch.o2, ch.o0), ch.vis);
                    }
                    float u = puct ? __fdiv_rn(__fmul_rn(__fmul_rn(cfg.c, __uint_as_float(ch.prior)), pterm), __fadd_rn(1.0f, ch.vis))
                                   : __fdiv_rn(pterm, __fsqrt_rn(ch.vis));
                    float value = __fadd_rn(q, u);
                    if (k == 0u || value > bval) { b = k; bval = value; bvis = ch.vis; bo0 = ch.o0; bo2 = ch.o2; bfc = ch.fc; bpk = ch.pk; }
                }
            }
        }
        g.cnt[CNT_SELECT_LEVELS] += 1u;
        g.cnt[CNT_CHILDREN_SCANNED] += nch;
        cur = cfc + b;
        cvis = bvis; cop0 = bo0; cop2 = bo2; cfc = bfc; cpk = bpk;
        c4::step(my, op, (int)((cpk >> 16) & 0xffu));
        if (++depth >= 64u) { g.err = DERR_DEPTH_OVERFLOW; return false; }
    }
    // ---- visit (mcts.rs:374-406): number the children of `cur`; auto-extend through only-children
    for (;;) {
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
        if (fc + n > c.cap) { g.err = DERR_ARENA_OVERFLOW; return false; }
        uint32_t* mw = meta_words(nodes, cur); // mark_visited (mcts.rs:399-400)
        mw[1] = fc;
        mw[3] = (cpk & 0xffffff00u) | n;
        g.nn = fc + n;
        g.cnt[CNT_EXPANSIONS] += 1u;
        g.cnt[CNT_CHILDREN_CREATED] += n;
        if (cfg.auto_extend && n == 1u) { // mcts.rs:404-405: recurse into the only child, no policy call
            int only = __ffs(lm) - 1;
            uint32_t s2 = (cs2 >> (2 * only)) & 3u;
            uint32_t osol = s2 == 1u ? c4::SOL_LOSE0 : (s2 == 2u ? c4::SOL_DRAW0 : 0u);
            cpk = (osol << 8) | ((uint32_t)only << 16);
            store_new_child(nodes, fc, 1.0f, cur, cpk);
            c4::step(my, op, only);
            cur = fc;
            if (++depth >= 64u) { g.err = DERR_DEPTH_OVERFLOW; return false; }
            if (osol) { backprop_onehot(c, g, cur, osol); return false; } // mcts.rs:377-379
            continue;
        }
        lf.id = cur; lf.fc = fc; lf.legal = lm; lf.csol2 = cs2; lf.any_solved = cs2 != 0u;
        return true;
    }
}

// The rest of visit() after Policy::eval (mcts.rs:384-397 child records, 409-423 stable softmax over
// the legal children in child order), then backprop.  logits[col] is used for legal columns only.
__device__ __forceinline__ void finish(const Ctx& c, Game& g, const Leaf& lf, bool uniform_prior, const float (&logits)[9],
                                       float o0, float o1, float o2) {
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
    const float up = __fdiv_rn(1.0f, (float)n); // all logits equal: exp(0) = 1 per child, total = n exactly
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
    backprop(c, g, lf.id, o0, o1, o2, lf.any_solved);
}

// mcts.rs:229-269 after the construction visit.  Cold path.
__device__ __noinline__ void add_root_noise(const Ctx& c, Game& g) {
    const syn_mcts_cfg& cfg = *c.cfg;
    if (cfg.noise_kind == SYN_NOISE_NONE) return;
    uint4* nodes = g.nodes;
    Rec r = load_rec(nodes, 0u);
    uint32_t nch = r.pk & 0xffu;
    if (nch < 2u) return;
    float w = cfg.noise_weight;
    float vals[9];
    for (int k = 0; k < 9; ++k) vals[k] = __fdiv_rn(1.0f, (float)nch);
    if (cfg.noise_kind == SYN_NOISE_DIRICHLET) {
        rng::Stream st;
        st.init(c.stream_seed(g, SYN_STREAM_NOISE), g.noise_pos);
        syn_dirichlet(st, cfg.noise_alpha, (int)nch, vals);
        g.noise_pos = (uint32_t)st.pos;
    }
    for (uint32_t k = 0; k < nch; ++k) {
        uint32_t* mw = meta_words(nodes, r.fc + k);
        float pr = __uint_as_float(mw[0]);
        pr = __fadd_rn(__fmul_rn(pr, __fsub_rn(1.0f, w)), __fmul_rn(w, vals[k]));
        mw[0] = __float_as_uint(pr);
    }
}

struct RootOut {
    float pi[9], visits[9];
    uint32_t child_sol[9];
    float q0, q1, q2;
    uint32_t root_sol, legal;
    int best_action;
};

// What the driver reads from a finished tree (mcts.rs:174-225, 273-306), by COLUMN.
__device__ __noinline__ void read_root(const Game& g, uint32_t action_selection, RootOut& r) {
    const uint4* nodes = g.nodes;
    Rec root = load_rec(nodes, 0u);
    uint32_t nch = root.pk & 0xffu, rsol = (root.pk >> 8) & 0xffu;
    for (int k = 0; k < 9; ++k) { r.pi[k] = 0.0f; r.visits[k] = 0.0f; r.child_sol[k] = 0u; }
    float total = 0.0f, b0 = 0.0f, b1 = 0.0f;
    int best = 0;
    uint32_t legal = 0u;
    for (uint32_t k = 0; k < nch; ++k) {
        Rec ch = load_rec(nodes, root.fc + k);
        uint32_t csol = (ch.pk >> 8) & 0xffu, act = (ch.pk >> 16) & 0xffu;
        float v; // target_policy (mcts.rs:174-211)
        if (root.vis == 1.0f) v = sol_kind(rsol) == SYN_KIND_WIN ? (sol_kind(csol) == SYN_KIND_LOSE ? 1.0f : 0.0f) : 1.0f;
        else v = ch.vis;
        total = __fadd_rn(total, v);
        float k0, k1; // best_action (mcts.rs:273-294): key (k0, k1), strict lexicographic >, first child incumbent
        uint32_t kind = sol_kind(csol);
        if (kind == SYN_KIND_WIN) { k0 = 0.0f; k1 = (float)(csol & 63u); }
        else if (kind == 0u) { k0 = 1.0f; k1 = action_selection == SYN_ACTION_Q ? -__fdiv_rn(__fsub_rn(ch.o2, ch.o0), ch.vis) : ch.vis; }
        else if (kind == SYN_KIND_DRAW) { k0 = 2.0f; k1 = -(float)(csol & 63u); }
        else { k0 = 3.0f; k1 = -(float)(csol & 63u); }
        if (k == 0u || k0 > b0 || (k0 == b0 && k1 > b1)) { b0 = k0; b1 = k1; best = (int)act; }
        legal |= 1u << act;
        // act is a runtime index: a tiny switch keeps the arrays out of local memory
#pragma unroll
        for (int col = 0; col < 9; ++col)
            if ((int)act == col) { r.pi[col] = v; r.visits[col] = ch.vis; r.child_sol[col] = csol; }
    }
#pragma unroll
    for (int col = 0; col < 9; ++col) r.pi[col] = __fdiv_rn(r.pi[col], total); // illegal columns: 0 / total = 0
    r.legal = legal;
    r.best_action = best;
    r.root_sol = rsol;
    if (rsol) { // target_q (mcts.rs:213-225)
        int idx = sol_index(rsol);
        r.q0 = idx == 0 ? 1.0f : 0.0f; r.q1 = idx == 1 ? 1.0f : 0.0f; r.q2 = idx == 2 ? 1.0f : 0.0f;
    } else {
        r.q0 = __fdiv_rn(root.o0, root.vis); r.q1 = __fdiv_rn(root.o1, root.vis); r.q2 = __fdiv_rn(root.o2, root.vis);
    }
}

// Ends the current move (alpha_zero.rs:246-267, 270-338): emit the row (or the search outputs),
// choose and play the action, and either start the next tree or close the game.  Cold path (once
// per tree).
__device__ __noinline__ void end_of_move(const KParams& p, const Ctx& c, Game& g) {
    const syn_rollout_cfg& cfg = p.cfg;
    RootOut r;
    read_root(g, cfg.action_selection, r);
    g.cnt[CNT_NODES] += g.nn;
    if (p.search_mode) {
        size_t i = g.gi;
        for (int k = 0; k < 9; ++k) {
            if (p.s_child_visits) p.s_child_visits[i * 9 + k] = r.visits[k];
            if (p.s_child_sol) p.s_child_sol[i * 9 + k] = (uint8_t)r.child_sol[k];
        }
        if (p.s_root_q) { p.s_root_q[i * 3 + 0] = r.q0; p.s_root_q[i * 3 + 1] = r.q1; p.s_root_q[i * 3 + 2] = r.q2; }
        if (p.s_root_sol) p.s_root_sol[i] = (uint8_t)r.root_sol;
        if (p.s_best) p.s_best[i] = (uint8_t)r.best_action;
        if (p.s_nodes) p.s_nodes[i] = g.nn;
        g.cnt[CNT_GAMES] += 1u;
        g.phase = PH_NEED_GAME;
        return;
    }
    size_t row = (size_t)g.gi * 63 + g.ply;
    for (int k = 0; k < 9; ++k) {
        p.row_pi[row * 9 + k] = r.pi[k];
        p.row_visits[row * 9 + k] = r.visits[k];
    }
    // sample_action (alpha_zero.rs:270-294)
    int action = r.best_action;
    uint32_t best_sol = 0u;
    for (int k = 0; k < 9; ++k)
        if (k == action) best_sol = r.child_sol[k];
    int mode = -1;
    if (g.ply < cfg.random_actions_until) mode = 0;
    else if (g.ply < cfg.sample_actions_until && (best_sol == 0u || !cfg.stop_games_when_solved)) mode = 1;
    if (mode >= 0) {
        uint32_t ap = g.apos;
        int a = sample_action_slow(c.stream_seed(g, SYN_STREAM_ACTION), &ap, mode, r.legal, r.pi);
        g.apos = ap;
        if (a < 0 || a > 8) { g.err = DERR_BAD_WEIGHTS; a = r.best_action; }
        action = a;
    }
    uint32_t solution = 0u; // mcts.solution(&action)
    for (int k = 0; k < 9; ++k)
        if (k == action) solution = r.child_sol[k];
    p.row_my[row] = g.my;
    p.row_op[row] = g.op;
    p.row_v[row * 3 + 0] = r.q0; p.row_v[row * 3 + 1] = r.q1; p.row_v[row * 3 + 2] = r.q2; // StateInfo::q
    p.row_action[row] = (uint8_t)action;
    p.row_nodes[row] = g.nn;
    uint32_t over = c4::step(g.my, g.op, action); // Outcome::from(reward(player)) when the game ended
    g.ply += 1;
    uint32_t fin = over ? over : (cfg.stop_games_when_solved ? solution : 0u);
    if (fin == 0u) { g.phase = PH_NEW_TREE; return; }
    // fill_state_info + store_rewards (alpha_zero.rs:296-338)
    uint32_t n = g.ply;
    uint32_t okind = 4u - sol_kind(fin); // solution.reversed(): the last mover's outcome
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t kind = okind;
        if (((n - 1u - k) & 1u) && kind != SYN_KIND_DRAW) kind = 4u - kind;
        size_t rr = (size_t)g.gi * 63 + k;
        float q0 = p.row_v[rr * 3 + 0], q1 = p.row_v[rr * 3 + 1], q2 = p.row_v[rr * 3 + 2];
        float z0 = kind == SYN_KIND_LOSE ? 1.0f : 0.0f, z1 = kind == SYN_KIND_DRAW ? 1.0f : 0.0f, z2 = kind == SYN_KIND_WIN ? 1.0f : 0.0f;
        float v0, v1, v2;
        if (cfg.value_target_kind == SYN_VALUE_Q) { v0 = q0; v1 = q1; v2 = q2; }
        else if (cfg.value_target_kind == SYN_VALUE_Z) { v0 = z0; v1 = z1; v2 = z2; }
        else if (cfg.value_target_kind == SYN_VALUE_QZ_AVERAGE) {
            float pp = cfg.vt_a, om = __fsub_rn(1.0f, pp);
            v0 = __fadd_rn(__fmul_rn(q0, pp), __fmul_rn(z0, om));
            v1 = __fadd_rn(__fmul_rn(q1, pp), __fmul_rn(z1, om));
            v2 = __fadd_rn(__fmul_rn(q2, pp), __fmul_rn(z2, om));
        } else {
            float tt = __fdiv_rn((float)(k + 1u), (float)n);
            float pp = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, tt), cfg.vt_a), __fmul_rn(tt, cfg.vt_b));
            float om = __fsub_rn(1.0f, pp);
            v0 = __fadd_rn(__fmul_rn(q0, om), __fmul_rn(z0, pp));
            v1 = __fadd_rn(__fmul_rn(q1, om), __fmul_rn(z1, pp));
            v2 = __fadd_rn(__fmul_rn(q2, om), __fmul_rn(z2, pp));
        }
        p.row_v[rr * 3 + 0] = v0; p.row_v[rr * 3 + 1] = v1; p.row_v[rr * 3 + 2] = v2;
    }
    p.game_len[g.gi] = n;
    g.cnt[CNT_ROWS] += n;
    g.cnt[CNT_GAMES] += 1u;
    g.phase = PH_NEED_GAME;
}

__device__ __forceinline__ void flush_counters(const KParams& p, Game& g) {
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) {
        if (g.cnt[i]) atomicAdd(p.counters + i, (unsigned long long)g.cnt[i]);
        g.cnt[i] = 0u;
    }
}

__device__ __forceinline__ void fail(const KParams& p, Game& g) {
    atomicCAS(p.error, 0, g.err);
    g.phase = PH_DONE;
}

// Runs the thread's state machine until a leaf needs Policy::eval (returns true; `lf` filled, the
// leaf's board in my/op) or no games are left (returns false, phase == PH_DONE).
__device__ __forceinline__ bool advance(const KParams& p, const Ctx& c, Game& g, Leaf& lf, uint64_t& my, uint64_t& op) {
    for (;;) {
        if (g.phase == PH_DONE) return false;
        if (g.phase == PH_NEED_GAME) {
            flush_counters(p, g);
            uint32_t gi = atomicAdd(p.next_game, 1u);
            if (gi >= p.num_games || *(volatile int*)p.error != 0) { g.phase = PH_DONE; return false; }
            g.gi = gi;
            g.ply = 0; g.apos = 0; g.fpu_pos = 0; g.noise_pos = 0;
            if (p.search_mode) {
                g.my = p.pos_my[gi];
                g.op = p.pos_op[gi];
                g.stream = p.pos_seed[gi];
            } else {
                g.my = 0; g.op = 0;
                g.stream = p.first_game + gi;
            }
            g.phase = PH_NEW_TREE;
        }
        if (g.phase == PH_NEW_TREE) { // MCTS::with_capacity (mcts.rs:123-137): fresh arena, root only
            g.nodes[0] = make_uint4(0u, 0u, 0u, 0u);
            g.nodes[1] = make_uint4(0u, 0u, 0u, 0u);
            g.nn = 1u;
            g.e_done = 0u;
            g.is_init = true;
            g.cnt[CNT_TREES] += 1u;
            g.phase = PH_EXPLORE;
        } else { // explore_n (mcts.rs:139-147): stop at num_explores or as soon as the root is solved
            uint32_t rpk = meta_words(g.nodes, 0u)[3];
            if (g.e_done >= p.cfg.num_explores || ((rpk >> 8) & 0xffu) != 0u) {
                end_of_move(p, c, g);
                if (g.err) { fail(p, g); return false; }
                continue;
            }
        }
        if (!g.is_init) g.cnt[CNT_EXPLORES] += 1u;
        my = g.my; op = g.op;
        bool need = descend(c, g, my, op, lf);
        if (g.err) { fail(p, g); return false; }
        if (need) return true;
        if (g.is_init) { add_root_noise(c, g); g.is_init = false; }
        else g.e_done += 1u;
    }
}

__device__ __forceinline__ void after_eval(const Ctx& c, Game& g) {
    if (g.is_init) { add_root_noise(c, g); g.is_init = false; }
    else g.e_done += 1u;
}

__device__ __forceinline__ void init_game(const KParams& p, Game& g, size_t slot) {
    g.nodes = p.nodes + 2 * slot * p.arena_nodes;
    g.nn = 1u;
    g.my = g.op = 0ull; g.stream = 0ull;
    g.gi = g.ply = g.e_done = g.apos = g.fpu_pos = g.noise_pos = 0u;
    g.phase = PH_NEED_GAME;
    g.is_init = false;
    g.err = 0;
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) g.cnt[i] = 0u;
}

} // namespace tpg
