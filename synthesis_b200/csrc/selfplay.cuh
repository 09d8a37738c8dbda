// selfplay.cuh — the persistent self-play kernel: every lane-group plays whole games.
//
// Replaces synthesis/src/alpha_zero.rs:181-338 of the reference (run_n_games, run_game,
// sample_action, fill_state_info, store_rewards) and the per-move driving of MCTS
// (with_capacity + explore_n + target_policy/target_q + best_action/solution).
//
// Groups pull game indices from one global counter (the reference splits games over
// num_workers+1 threads up front, alpha_zero.rs:132-154), play them to the end and write one
// experience row per ply at rows[game*63 + ply]; a compaction kernel (engine.cu) then lays the
// rows out densely in game order, which is the order ReplayBuffer::extend produces.
//
// Leaf evaluation:
//   rollout mode — RolloutPolicy inline in the group, no synchronisation with anyone;
//   NN mode      — each group parks its leaf's features in the CTA's shared activation tile, the
//                  CTA runs ONE batched Connect4Net forward for all of its groups, and every
//                  group then finishes its explore.  One pending leaf per tree, so per-tree
//                  semantics stay strictly sequential (no virtual loss).
#pragma once
#include "mlp.cuh"
#include "mlp_tc.cuh"
#include "tree.cuh"

namespace eng {

struct KParams {
    syn_rollout_cfg cfg;
    uint64_t seed, first_game;
    uint32_t num_games;
    uint32_t search_mode; // 1: one tree per given position, no game loop (syn_engine_search)
    uint32_t no_reductions; // tpg2.cuh: 1 = backprop by load / add / store only (SYN_TPG_NO_RED=1; the parity suite runs both)
    // seating (tp2::seat_of): CTA b plays seats_q + (b < seats_rem) games, dealt round-robin to teams_used teams of at most
    // per_team seats each (rollout kernel: one "team" = the CTA), so that few games still occupy every SM
    uint32_t teams_used, per_team, seats_q, seats_rem;
    uint32_t arena_nodes;
    uint4* nodes; // tree arenas: arena_nodes 32-byte records per game slot (tree.cuh)
    unsigned int* next_game;
    uint32_t* slot_state; // tpg2.cuh: 8 words of cold per-game state per arena slot
    // experience rows, [num_games][63]
    uint64_t* row_my;
    uint64_t* row_op;
    float* row_pi;     // [..][9]
    float* row_v;      // [..][3]
    uint8_t* row_action;
    uint32_t* row_nodes;
    float* row_visits; // [..][9]
    uint32_t* game_len;
    // search mode inputs / outputs, [num_games]
    const uint64_t* pos_my;
    const uint64_t* pos_op;
    const uint64_t* pos_seed;
    float* s_child_visits; // [..][9]
    uint8_t* s_child_sol;  // [..][9]
    float* s_root_q;       // [..][3]
    uint8_t* s_root_sol;
    uint8_t* s_best;
    uint32_t* s_nodes;
    unsigned long long* counters; // [CNT_N]
    int* error;
    const float* weights; // device blob, NN mode
    const uint8_t* weight_image; // mlptc image (fp16 UMMA layout), NN mode on tensor cores
    const uint8_t* weight_image_lo; // mlp_split.cuh: the fp16 remainders W - fp16(W) in the same layout
    float mlp_bias[mlptc::BIAS_FLOATS]; // the image's padded fp32 biases again, in the parameter space: constant-bank operands (mlp_team.cuh forward_cb)
    uint32_t* fpu_state;                // [max_games][tp2::FS_WORDS]: per-slot cache of the Normal-FPU stream (thread-per-game kernels); last member:
                                        // the offsets of everything above are what the kernels were tuned with
};

enum Phase { PH_NEED_GAME = 0, PH_NEW_TREE = 1, PH_EXPLORE = 2, PH_DONE = 3 };

// alpha_zero.rs:280-287 on the game's action stream.  One lane.  mode 0: uniform legal move
// (gen_range(0..count as u8)); mode 1: WeightedIndex::new(pi).sample.  Returns -1 on a weight error.
__device__ __noinline__ int sample_action_slow(uint64_t aseed, uint32_t* apos, int mode, uint32_t legal, const float* pi) {
    rng::Stream st;
    st.init(aseed, *apos);
    int action;
    if (mode == 0) {
        uint32_t n = (uint32_t)__popc(legal);
        uint32_t i = st.gen_range_u8(n);
        uint32_t m = legal;
        for (uint32_t k = 0; k < i; ++k) m &= m - 1u;
        action = __ffs(m) - 1;
    } else {
        // rand 0.8 WeightedIndex<f32>: cumulative sums of all but the last weight; Uniform(0,total)
        float cum[8];
        float total = pi[0];
        bool ok = total >= 0.0f;
        for (int i = 1; i < 9; ++i) {
            ok = ok && (pi[i] >= 0.0f);
            cum[i - 1] = total;
            total = __fadd_rn(total, pi[i]);
        }
        if (!ok || !(total > 0.0f)) {
            action = -1;
        } else {
            const float max_rand = 1.0f - 1.1920929e-7f;
            float scale = total; // high - low with low = 0
            for (;;) {
                float top = __fadd_rn(__fmul_rn(scale, max_rand), 0.0f);
                if (!(top >= total)) break;
                scale = __uint_as_float(__float_as_uint(scale) - 1u);
            }
            float x = __fadd_rn(__fmul_rn(st.next_f32_01(), scale), 0.0f);
            action = 0;
            for (int i = 0; i < 8; ++i) {
                if (cum[i] <= x) ++action;
                else break;
            }
        }
    }
    *apos = (uint32_t)st.pos;
    return action;
}

template <int GL, bool NN>
struct GroupState {
    int phase;
    bool is_init;      // the explore in flight is the construction visit of MCTS::with_capacity
    uint32_t gi;       // index of the game within this call
    uint64_t my, op;   // root position of the current tree
    uint32_t ply;
    uint32_t e_done;
    uint64_t aseed;
    uint32_t apos;
};

// Ends the current move: read the tree, emit the row (or the search outputs), choose and play the
// action, and either start the next tree or close the game.
template <int GL, bool NN>
__device__ __forceinline__ void end_of_move(const Grp<GL>& g, const KParams& p, Tree<GL>& t, GroupState<GL, NN>& st) {
    const syn_rollout_cfg& cfg = p.cfg;
    RootReadout r;
    read_root(g, t, cfg.action_selection, r);
    t.cnt[CNT_NODES] += t.nn;
    if (p.search_mode) {
        size_t i = st.gi;
        if (g.gl < 9) {
            if (p.s_child_visits) p.s_child_visits[i * 9 + g.gl] = r.visits;
            if (p.s_child_sol) p.s_child_sol[i * 9 + g.gl] = (uint8_t)r.child_sol;
        }
        if (g.gl == 0) {
            if (p.s_root_q) { p.s_root_q[i * 3 + 0] = r.q0; p.s_root_q[i * 3 + 1] = r.q1; p.s_root_q[i * 3 + 2] = r.q2; }
            if (p.s_root_sol) p.s_root_sol[i] = (uint8_t)r.root_sol;
            if (p.s_best) p.s_best[i] = (uint8_t)r.best_action;
            if (p.s_nodes) p.s_nodes[i] = t.nn;
        }
        t.cnt[CNT_GAMES] += 1u;
        st.phase = PH_NEED_GAME;
        return;
    }
    size_t row = (size_t)st.gi * 63 + st.ply;
    if (g.gl < 9) {
        p.row_pi[row * 9 + g.gl] = r.pi;
        p.row_visits[row * 9 + g.gl] = r.visits;
    }
    // sample_action (alpha_zero.rs:270-294)
    uint32_t best_sol = g.shfl(r.child_sol, r.best_action);
    int action = r.best_action;
    int mode = -1;
    if (st.ply < cfg.random_actions_until) mode = 0;
    else if (st.ply < cfg.sample_actions_until && (best_sol == 0u || !cfg.stop_games_when_solved)) mode = 1;
    if (mode >= 0) {
        float pis[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) pis[k] = g.shfl(r.pi, k);
        int a = 0;
        uint32_t np = 0;
        if (g.gl == 0) {
            uint32_t ap = st.apos;
            a = sample_action_slow(st.aseed, &ap, mode, r.legal, pis);
            np = ap;
        }
        action = (int)g.shfl((uint32_t)a, 0);
        st.apos = g.shfl(np, 0);
        if (action < 0 || action > 8) { t.err = DERR_BAD_WEIGHTS; action = r.best_action; }
    }
    uint32_t solution = g.shfl(r.child_sol, action); // mcts.solution(&action)
    if (g.gl == 0) {
        p.row_my[row] = st.my;
        p.row_op[row] = st.op;
        p.row_v[row * 3 + 0] = r.q0; p.row_v[row * 3 + 1] = r.q1; p.row_v[row * 3 + 2] = r.q2; // StateInfo::q
        p.row_action[row] = (uint8_t)action;
        p.row_nodes[row] = t.nn;
    }
    uint32_t over = c4::step(st.my, st.op, action); // Outcome::from(reward(player)) when the game ended
    st.ply += 1;
    uint32_t fin = over ? over : (cfg.stop_games_when_solved ? solution : 0u);
    if (fin == 0u) { st.phase = PH_NEW_TREE; return; }
    // fill_state_info + store_rewards (alpha_zero.rs:296-338)
    g.sync();
    uint32_t n = st.ply;
    uint32_t okind = 4u - sol_kind(fin); // solution.reversed(): the last mover's outcome
    for (uint32_t k = g.gl; k < n; k += GL) {
        uint32_t kind = okind;
        if (((n - 1u - k) & 1u) && kind != SYN_KIND_DRAW) kind = 4u - kind;
        size_t rr = (size_t)st.gi * 63 + k;
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
    if (g.gl == 0) p.game_len[st.gi] = n;
    t.cnt[CNT_ROWS] += n;
    t.cnt[CNT_GAMES] += 1u;
    st.phase = PH_NEED_GAME;
}

template <int GL>
__device__ __forceinline__ void flush_counters(const Grp<GL>& g, const KParams& p, Tree<GL>& t) {
    if (g.gl == 0) {
#pragma unroll
        for (int i = 0; i < CNT_N; ++i)
            if (t.cnt[i]) atomicAdd(p.counters + i, (unsigned long long)t.cnt[i]);
    }
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) t.cnt[i] = 0u;
}

// Seating of the lane-group kernels: CTA b runs seats_q + (b < seats_rem) groups, so that few games are dealt over all SMs
// (and both CTAs of an SM) instead of filling the first CTAs — with network leaves a round ends at a CTA barrier, and the
// fewer groups share it the less each waits for the slowest descent.  Idle groups still join the barriers.
__device__ __forceinline__ bool lg_seated(const KParams& p, int grp) { return (uint32_t)grp < p.seats_q + (blockIdx.x < p.seats_rem ? 1u : 0u); }

#ifdef SYN_LG_PROF
// Phase clocks of the lane-group kernels (a -DSYN_LG_PROF build only; syn_engine_debug_counters): one set per group.
struct LgProf {
    long long adv = 0, wait = 0, leaf = 0, fin = 0, rounds = 0, t_start;
    __device__ __forceinline__ LgProf() { t_start = clock64(); }
    template <int GL>
    __device__ __forceinline__ void flush(const Grp<GL>& g, const KParams& p, Tree<GL>& t) {
        if (g.gl != 0) return;
        atomicAdd(p.counters + DBG_T_ADVANCE, (unsigned long long)adv);
        atomicAdd(p.counters + DBG_T_TEAMWAIT, (unsigned long long)wait);
        atomicAdd(p.counters + DBG_T_MLP, (unsigned long long)leaf);
        atomicAdd(p.counters + DBG_T_FINISH, (unsigned long long)fin);
        atomicAdd(p.counters + DBG_ROUNDS, (unsigned long long)rounds);
        atomicAdd(p.counters + DBG_LEAVES, 1ull); // groups
        atomicAdd(p.counters + DBG_T_TOTAL, (unsigned long long)(clock64() - t_start));
        atomicAdd(p.counters + DBG_X_SELECT, (unsigned long long)t.pt[0]);
        atomicAdd(p.counters + DBG_X_EXPAND, (unsigned long long)t.pt[1]);
        atomicAdd(p.counters + DBG_X_EOM, (unsigned long long)t.pt[2]);
        atomicAdd(p.counters + DBG_X_BACKPROP, (unsigned long long)t.pt[3]);
    }
};
#define LGP_INIT(t) LgProf lgp; (t).pt[0] = (t).pt[1] = (t).pt[2] = (t).pt[3] = 0
#define LGP_MARK(field, since) do { long long _n = clock64(); lgp.field += _n - (since); (since) = _n; } while (0)
#define LGP_FLUSH(g, p, t) lgp.flush(g, p, t)
#else
#define LGP_INIT(t) ((void)0)
#define LGP_MARK(field, since) ((void)(since))
#define LGP_FLUSH(g, p, t) ((void)0)
#endif

// Runs the group's state machine until a leaf needs Policy::eval (returns true, `pend` filled) or
// no games are left (returns false, phase == PH_DONE).
template <int GL, bool NN>
__device__ __forceinline__ bool advance(const Grp<GL>& g, const KParams& p, Tree<GL>& t, GroupState<GL, NN>& st,
                                        RolloutRng<GL>& rr, uint32_t* rr_smem, Pending& pend) {
    for (;;) {
        if (st.phase == PH_DONE) return false;
        if (st.phase == PH_NEED_GAME) {
            flush_counters(g, p, t);
            uint32_t gi = 0;
            if (g.gl == 0) gi = atomicAdd(p.next_game, 1u);
            gi = g.shfl(gi, 0);
            if (gi >= p.num_games || *(volatile int*)p.error != 0) { st.phase = PH_DONE; return false; }
            st.gi = gi;
            st.ply = 0;
            st.apos = 0;
            t.fpu_pos = 0;
            t.noise_pos = 0;
            uint64_t rseed;
            if (p.search_mode) {
                st.my = p.pos_my[gi];
                st.op = p.pos_op[gi];
                rseed = p.pos_seed[gi];
                st.aseed = 0;
                t.noise_seed = rseed ^ (1ull << 63);
                t.fpu_seed = (rseed ^ (1ull << 63)) + 1ull;
            } else {
                uint64_t G = p.first_game + gi;
                st.my = 0; st.op = 0;
                rseed = syn_stream_seed(p.seed, G, SYN_STREAM_ROLLOUT);
                st.aseed = syn_stream_seed(p.seed, G, SYN_STREAM_ACTION);
                t.noise_seed = syn_stream_seed(p.seed, G, SYN_STREAM_NOISE);
                t.fpu_seed = syn_stream_seed(p.seed, G, SYN_STREAM_FPU);
            }
            if (!NN) rr.init(g, rseed, rr_smem);
            if (p.cfg.mcts.fpu_kind == SYN_FPU_NORMAL) fpu_stream_begin(g, t);
            st.phase = PH_NEW_TREE;
        }
        if (st.phase == PH_NEW_TREE) { // MCTS::with_capacity (mcts.rs:123-137): fresh arena, root only
            g.sync();
            if (g.gl == 0) {
                t.meta[0] = make_uint4(0u, 0u, 0u, 0u);
                t.stat[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            t.nn = 1u;
            st.e_done = 0u;
            st.is_init = true;
            t.cnt[CNT_TREES] += 1u;
            st.phase = PH_EXPLORE;
            g.sync();
        } else { // explore_n (mcts.rs:139-147): stop at num_explores or as soon as the root is solved
            uint32_t rpk = t.meta[0].w;
            if (st.e_done >= p.cfg.num_explores || ((rpk >> 8) & 0xffu) != 0u) {
                const long long lgp_eom = LGP_NOW();
                end_of_move(g, p, t, st);
                LGP_ADD(t, 2, lgp_eom);
                if (t.err) {
                    if (g.gl == 0) atomicCAS(p.error, 0, t.err);
                    st.phase = PH_DONE;
                    return false;
                }
                continue;
            }
        }
        if (!st.is_init) t.cnt[CNT_EXPLORES] += 1u;
        bool need = explore_descend(g, t, st.my, st.op, pend);
        if (t.err) {
            if (g.gl == 0) atomicCAS(p.error, 0, t.err);
            st.phase = PH_DONE;
            return false;
        }
        if (need) return true;
        if (st.is_init) { add_root_noise(g, t); st.is_init = false; }
        else st.e_done += 1u;
    }
}

template <int GL, bool NN>
__device__ __forceinline__ void after_eval(const Grp<GL>& g, Tree<GL>& t, GroupState<GL, NN>& st) {
    if (st.is_init) { add_root_noise(g, t); st.is_init = false; }
    else st.e_done += 1u;
}

// ------------------------------------------------------------------ rollout-mode kernel
template <int GL, int THREADS>
__global__ void __launch_bounds__(THREADS) selfplay_rollout_kernel(const __grid_constant__ KParams p) {
    constexpr int GPB = THREADS / GL;
    __shared__ uint32_t s_path[GPB][64];
    __shared__ uint32_t s_rng[GPB][4 * GL];
    __shared__ uint32_t s_fpu[GPB][FPU_SM_WORDS];
    Grp<GL> g;
    const int grp = threadIdx.x / GL;
    const size_t slot = (size_t)blockIdx.x * GPB + grp;
    Tree<GL> t;
    t.stat.base = t.meta.base = p.nodes + 2 * slot * p.arena_nodes;
    t.path = s_path[grp];
    t.fpu_sm = s_fpu[grp];
    t.cap = p.arena_nodes;
    t.cfg = &p.cfg.mcts;
    t.err = 0;
    t.nn = 1;
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) t.cnt[i] = 0u;
    GroupState<GL, false> st;
    st.phase = lg_seated(p, grp) ? PH_NEED_GAME : PH_DONE;
    st.is_init = false;
    RolloutRng<GL> rr;
    rr.buf = s_rng[grp]; rr.kA = rr.kB = 0; rr.pos = 0;
    Pending pend;
    LGP_INIT(t);
    long long lgp_t = LGP_NOW();
    while (advance<GL, false>(g, p, t, st, rr, s_rng[grp], pend)) {
        LGP_MARK(adv, lgp_t);
        uint32_t plies = 0;
        int idx = rollout(g, pend.my, pend.op, rr, plies);
        t.cnt[CNT_ROLLOUT_PLIES] += plies;
        LGP_MARK(leaf, lgp_t);
        explore_finish(g, t, pend, true, 0.0f, idx == 0 ? 1.0f : 0.0f, idx == 1 ? 1.0f : 0.0f, idx == 2 ? 1.0f : 0.0f);
        after_eval(g, t, st);
        LGP_MARK(fin, lgp_t);
#ifdef SYN_LG_PROF
        ++lgp.rounds;
#endif
    }
    flush_counters(g, p, t);
    LGP_FLUSH(g, p, t);
}

// ------------------------------------------------------------------ NN-mode kernel (fp32 CUDA-core MLP)
// Shared memory: transposed weights | activation tile A | activation tile B | paths.
template <int GL, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) selfplay_nn_kernel(const __grid_constant__ KParams p) {
    constexpr int GPB = THREADS / GL; // groups (= leaves per batched forward) per CTA
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;
    float* xa = sw + mlp::WEIGHT_FLOATS;
    float* xb = xa + GPB * mlp::XS;
    uint32_t* s_path = reinterpret_cast<uint32_t*>(xb + GPB * mlp::XS);
    mlp::load_weights_transposed(sw, p.weights, threadIdx.x, THREADS);

    Grp<GL> g;
    const int grp = threadIdx.x / GL;
    const size_t slot = (size_t)blockIdx.x * GPB + grp;
    Tree<GL> t;
    t.stat.base = t.meta.base = p.nodes + 2 * slot * p.arena_nodes;
    t.path = s_path + grp * 64;
    t.cap = p.arena_nodes;
    t.cfg = &p.cfg.mcts;
    t.err = 0;
    t.nn = 1;
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) t.cnt[i] = 0u;
    GroupState<GL, true> st;
    st.phase = lg_seated(p, grp) ? PH_NEED_GAME : PH_DONE;
    st.is_init = false;
    RolloutRng<GL> rr;
    rr.buf = nullptr; rr.kA = rr.kB = 0; rr.pos = 0;
    Pending pend;
    for (;;) {
        bool need = advance<GL, true>(g, p, t, st, rr, nullptr, pend);
        if (need) { // Game::features of the leaf straight from its bitboards into the activation tile
            for (int i = g.gl; i < 64; i += GL) xa[grp * mlp::XS + i] = i < 63 ? c4::feature(pend.my, pend.op, i) : 0.0f;
        }
        if (!__syncthreads_or(need ? 1 : 0)) break; // no group of this CTA has work left
        mlp::forward<GPB, THREADS>(sw, xa, xb, threadIdx.x);
        if (need) {
            const float* y = xb + grp * mlp::XS;
            float logit = g.gl < 9 ? y[g.gl] : 0.0f;
            // value.softmax(-1) (study-connect4/src/policies.rs:54-56)
            float y0 = y[9], y1 = y[10], y2 = y[11];
            float m = fmaxf(y0, fmaxf(y1, y2));
            float e0 = syn_expf(__fsub_rn(y0, m)), e1 = syn_expf(__fsub_rn(y1, m)), e2 = syn_expf(__fsub_rn(y2, m));
            float tot = __fadd_rn(__fadd_rn(e0, e1), e2);
            explore_finish(g, t, pend, false, logit, __fdiv_rn(e0, tot), __fdiv_rn(e1, tot), __fdiv_rn(e2, tot));
            after_eval(g, t, st);
        }
        __syncthreads(); // xb is overwritten by the next forward's first layer
    }
    flush_counters(g, p, t);
}

// ------------------------------------------------------------------ NN-mode kernel, tensor-core MLP (tcgen05 + TMEM)
// Shared memory: mlptc::Smem (weight image, two activation tiles, outputs, barriers) | paths.
// Group i of the CTA owns tile row mlptc::row_of_slot(i); the whole CTA runs one five-GEMM chain per
// round for all of its groups' leaves.
template <int GL, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS <= 256 ? 2 : 1) selfplay_nn_tc_kernel(const __grid_constant__ KParams p) {
    constexpr int GPB = THREADS / GL;
    static_assert(GPB <= 128 && GPB % 4 == 0, "one tile row per group");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    mlptc::Smem& ms = *reinterpret_cast<mlptc::Smem*>(smem_raw);
    uint32_t* s_path = reinterpret_cast<uint32_t*>(smem_raw + sizeof(mlptc::Smem));
    mlptc::setup(ms, p.weight_image);

    Grp<GL> g;
    const int grp = threadIdx.x / GL;
    const int row = mlptc::row_of_slot(grp);
    const int row_of_grp = row;
    const size_t slot = (size_t)blockIdx.x * GPB + grp;
    Tree<GL> t;
    t.stat.base = t.meta.base = p.nodes + 2 * slot * p.arena_nodes;
    t.path = s_path + grp * 64;
    t.cap = p.arena_nodes;
    t.cfg = &p.cfg.mcts;
    t.err = 0;
    t.nn = 1;
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) t.cnt[i] = 0u;
    GroupState<GL, true> st;
    st.phase = lg_seated(p, grp) ? PH_NEED_GAME : PH_DONE;
    st.is_init = false;
    RolloutRng<GL> rr;
    rr.buf = nullptr; rr.kA = rr.kB = 0; rr.pos = 0;
    Pending pend;
    uint32_t mma_phase = 0;
    LGP_INIT(t);
    long long lgp_t = LGP_NOW();
    for (;;) {
        bool need = advance<GL, true>(g, p, t, st, rr, nullptr, pend);
        if (need) { // Game::features of the leaf, fp16, straight into the A tile (K-major UMMA layout, permuted K)
            for (int k = g.gl; k < 80; k += GL) {
                int col = k >> 3, row = k & 7;
                float f = (row < 7 && col < 9) ? c4::feature(pend.my, pend.op, row * 9 + col) : 0.0f;
                *reinterpret_cast<__half*>(ms.a + mlptc::a_off(row_of_grp, k)) = __float2half_rn(f);
            }
        }
        LGP_MARK(adv, lgp_t);
        if (!__syncthreads_or(need ? 1 : 0)) break;
        LGP_MARK(wait, lgp_t);
        mlptc::forward<THREADS / 32>(ms, mma_phase, GPB / 4);
        LGP_MARK(leaf, lgp_t);
#ifdef SYN_LG_PROF
        ++lgp.rounds;
#endif
        if (need) {
            const float* y = ms.y[row];
            float logit = g.gl < 9 ? y[g.gl] : 0.0f;
            float y0 = y[9], y1 = y[10], y2 = y[11];
            float m = fmaxf(y0, fmaxf(y1, y2));
            float e0 = syn_expf(__fsub_rn(y0, m)), e1 = syn_expf(__fsub_rn(y1, m)), e2 = syn_expf(__fsub_rn(y2, m));
            float tot = __fadd_rn(__fadd_rn(e0, e1), e2);
            explore_finish(g, t, pend, false, logit, __fdiv_rn(e0, tot), __fdiv_rn(e1, tot), __fdiv_rn(e2, tot));
            after_eval(g, t, st);
        }
        LGP_MARK(fin, lgp_t);
        // no trailing barrier needed: the next forward() starts with a CTA barrier before any MMA, and
        // y / the A tile are only rewritten after that barrier (the tile row by this thread's own group, y by the epilogue)
    }
    flush_counters(g, p, t);
    LGP_FLUSH(g, p, t);
    mlptc::teardown(ms);
}

// Batched Policy::eval on tensor cores: 128 positions per tile, row = position.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) eval_tc_kernel(const uint8_t* __restrict__ weight_image, const uint64_t* __restrict__ my_bb,
                                                              const uint64_t* __restrict__ op_bb, uint32_t n, float* __restrict__ logits,
                                                              float* __restrict__ probs) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    mlptc::Smem& ms = *reinterpret_cast<mlptc::Smem*>(smem_raw);
    mlptc::setup(ms, weight_image);
    uint32_t mma_phase = 0;
    for (uint32_t base = blockIdx.x * 128u; base < n; base += gridDim.x * 128u) {
        for (int e = threadIdx.x; e < 128 * 80; e += THREADS) {
            int r = e / 80, k = e - r * 80;
            int col = k >> 3, row = k & 7;
            uint32_t idx = base + r;
            float f = 0.0f;
            if (idx < n && row < 7 && col < 9) f = c4::feature(my_bb[idx], op_bb[idx], row * 9 + col);
            *reinterpret_cast<__half*>(ms.a + mlptc::a_off(r, k)) = __float2half_rn(f);
        }
        mlptc::forward<THREADS / 32>(ms, mma_phase, 32);
        if (threadIdx.x < 128) {
            uint32_t idx = base + threadIdx.x;
            if (idx < n) {
                const float* y = ms.y[threadIdx.x];
                for (int k = 0; k < 9; ++k) logits[(size_t)idx * 9 + k] = y[k];
                float m = fmaxf(y[9], fmaxf(y[10], y[11]));
                float e0 = syn_expf(__fsub_rn(y[9], m)), e1 = syn_expf(__fsub_rn(y[10], m)), e2 = syn_expf(__fsub_rn(y[11], m));
                float tot = __fadd_rn(__fadd_rn(e0, e1), e2);
                probs[(size_t)idx * 3 + 0] = __fdiv_rn(e0, tot);
                probs[(size_t)idx * 3 + 1] = __fdiv_rn(e1, tot);
                probs[(size_t)idx * 3 + 2] = __fdiv_rn(e2, tot);
            }
        }
        __syncthreads();
    }
    mlptc::teardown(ms);
}

// ------------------------------------------------------------------ batched Policy::eval (syn_engine_eval)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) eval_kernel(const float* __restrict__ weights, const uint64_t* __restrict__ my_bb,
                                                           const uint64_t* __restrict__ op_bb, uint32_t n, float* __restrict__ logits,
                                                           float* __restrict__ probs) {
    constexpr int R = 32;
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;
    float* xa = sw + mlp::WEIGHT_FLOATS;
    float* xb = xa + R * mlp::XS;
    mlp::load_weights_transposed(sw, weights, threadIdx.x, THREADS);
    for (uint32_t base = blockIdx.x * R; base < n; base += gridDim.x * R) {
        for (int e = threadIdx.x; e < R * 64; e += THREADS) {
            int row = e >> 6, i = e & 63;
            uint32_t idx = base + row;
            float v = 0.0f;
            if (idx < n && i < 63) v = c4::feature(my_bb[idx], op_bb[idx], i);
            xa[row * mlp::XS + i] = v;
        }
        __syncthreads();
        mlp::forward<R, THREADS>(sw, xa, xb, threadIdx.x);
        if (threadIdx.x < R) {
            uint32_t idx = base + threadIdx.x;
            if (idx < n) {
                const float* y = xb + threadIdx.x * mlp::XS;
                for (int k = 0; k < 9; ++k) logits[(size_t)idx * 9 + k] = y[k];
                float m = fmaxf(y[9], fmaxf(y[10], y[11]));
                float e0 = syn_expf(__fsub_rn(y[9], m)), e1 = syn_expf(__fsub_rn(y[10], m)), e2 = syn_expf(__fsub_rn(y[11], m));
                float tot = __fadd_rn(__fadd_rn(e0, e1), e2);
                probs[(size_t)idx * 3 + 0] = __fdiv_rn(e0, tot);
                probs[(size_t)idx * 3 + 1] = __fdiv_rn(e1, tot);
                probs[(size_t)idx * 3 + 2] = __fdiv_rn(e2, tot);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ game rules on move lists (syn_engine_play)
// One warp per game: lane c owns column c for the legal-move ballot and the feature plane.
__global__ void play_kernel(const uint8_t* __restrict__ moves, const uint32_t* __restrict__ n_moves, uint32_t stride, uint32_t n_games,
                            uint64_t* my_out, uint64_t* op_out, uint8_t* height_out, uint8_t* legal_lo, uint8_t* legal_hi,
                            uint8_t* status_out, float* features) {
    uint32_t gidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (gidx >= n_games) return;
    uint64_t my = 0, op = 0;
    uint32_t status = 0;
    uint32_t nm = n_moves[gidx];
    bool illegal = false;
    for (uint32_t k = 0; k < nm; ++k) {
        int col = moves[(size_t)gidx * stride + k];
        uint64_t occ = my | op;
        bool has_room = lane < 9 && ((occ >> (7 * lane + 6)) & 1ull) == 0ull; // warp-wide legal-move detection
        unsigned lm = __ballot_sync(0xffffffffu, has_room);
        if (col >= 9 || !((lm >> col) & 1u)) { illegal = true; break; }
        uint32_t over = c4::step(my, op, col);
        status = over ? (1u | (over == c4::SOL_LOSE0 ? 2u : 0u)) : 0u;
    }
    uint64_t occ = my | op;
    bool has_room = lane < 9 && ((occ >> (7 * lane + 6)) & 1ull) == 0ull;
    unsigned lm = __ballot_sync(0xffffffffu, has_room);
    if (lane == 0) {
        my_out[gidx] = my;
        op_out[gidx] = op;
        legal_lo[gidx] = (uint8_t)(lm & 0xffu);
        legal_hi[gidx] = (uint8_t)((lm >> 8) & 1u);
        status_out[gidx] = illegal ? 255 : (uint8_t)status;
    }
    if (lane < 9) height_out[(size_t)gidx * 9 + lane] = (uint8_t)c4::height(occ, lane);
    if (features)
        for (int i = lane; i < 63; i += 32) features[(size_t)gidx * 63 + i] = c4::feature(my, op, i);
}

// ------------------------------------------------------------------ row compaction (ReplayBuffer layout)
struct CompactParams {
    uint32_t num_games;
    uint64_t first_game;
    const uint32_t* game_len;
    const uint64_t* row_off; // exclusive prefix sum of game_len
    const uint64_t* row_my;
    const uint64_t* row_op;
    const float* row_pi;
    const float* row_v;
    const uint8_t* row_action;
    const uint32_t* row_nodes;
    const float* row_visits;
    // dense outputs (any may be null)
    uint64_t* game_ids;
    uint64_t* my_bb;
    uint64_t* op_bb;
    uint8_t* height;
    uint8_t* player;
    float* states;
    float* pis;
    float* vs;
    uint8_t* t_action;
    uint32_t* t_nodes;
    float* t_visits;
};

// One warp per (game, ply) row; lanes spread over the 63 features.
__global__ void compact_kernel(const __grid_constant__ CompactParams c) {
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; // 64-bit: 2^32 threads are 2.1 M games
    int lane = threadIdx.x & 31;
    const uint64_t gi64 = w / 63u;
    if (gi64 >= c.num_games) return;
    const uint32_t gi = (uint32_t)gi64, k = (uint32_t)(w - gi64 * 63u);
    if (k >= c.game_len[gi]) return;
    size_t src = (size_t)gi * 63 + k, dst = c.row_off[gi] + k;
    uint64_t my = c.row_my[src], op = c.row_op[src];
    uint64_t occ = my | op;
    if (lane == 0) {
        if (c.game_ids) c.game_ids[dst] = c.first_game + gi + 1ull; // 1-based like ReplayBuffer::new_game (data.rs:128-130)
        if (c.my_bb) c.my_bb[dst] = my;
        if (c.op_bb) c.op_bb[dst] = op;
        if (c.player) c.player[dst] = (uint8_t)(__popcll(occ) & 1);
        if (c.t_action) c.t_action[dst] = c.row_action[src];
        if (c.t_nodes) c.t_nodes[dst] = c.row_nodes[src];
    }
    if (lane < 9) {
        if (c.height) c.height[dst * 9 + lane] = (uint8_t)c4::height(occ, lane);
        if (c.pis) c.pis[dst * 9 + lane] = c.row_pi[src * 9 + lane];
        if (c.t_visits) c.t_visits[dst * 9 + lane] = c.row_visits[src * 9 + lane];
    }
    if (lane < 3 && c.vs) c.vs[dst * 3 + lane] = c.row_v[src * 3 + lane];
    if (c.states)
        for (int i = lane; i < 63; i += 32) c.states[dst * 63 + i] = c4::feature(my, op, i);
}

// What a ReplayBuffer row holds beyond (game id, bitboards, pi, v) is a function of the bitboards: Connect4::height and
// ::player (connect4.rs:108-114) and Game::features (connect4.rs:237-258).  Rows travel between GPUs as those 72 bytes;
// the receiving rank rebuilds the rest here (a warp per row).
__global__ void expand_rows_kernel(uint64_t n_rows, const uint64_t* __restrict__ my_bb, const uint64_t* __restrict__ op_bb, uint8_t* height, uint8_t* player,
                                   float* states) {
    const uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const uint64_t my = my_bb[row], op = op_bb[row], occ = my | op;
    if (lane == 0 && player) player[row] = (uint8_t)(__popcll(occ) & 1);
    if (lane < 9 && height) height[row * 9 + lane] = (uint8_t)c4::height(occ, lane);
    if (states)
        for (int i = lane; i < 63; i += 32) states[row * 63 + i] = c4::feature(my, op, i);
}

} // namespace eng
