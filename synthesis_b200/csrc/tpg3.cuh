// tpg3.cuh — the round-synchronous thread-per-game kernel of tpg2.cuh on SPLIT node records.
//
// Replaces synthesis/src/mcts.rs:29-489 (Node, select_best_child, visit, backprop) of the
// reference; schedule, state machine and cold paths are tpg2.cuh's.  What changes is the data
// layout in HBM, because ncu put tpg2 at 2.2 TB/s of DRAM traffic in 32-byte random sectors
// (2.3 KB per explore, 1.9x the algorithmic bytes): select_best_child reads every child of every
// node on the path, but needs only 16 of a node's 32 bytes.
//
// A game's arena of `cap` nodes is two arrays:
//   hot[cap]  (16 B)  x = num_visits (u24, exact: visits are integers < 2^24) | solution << 24
//                     y = outcome_probs[2] - outcome_probs[0] (f32, the numerator of q(), mcts.rs:42-44)
//                     z = action_prob (f32)
//                     w = first_child (u24) | num_children << 24 | action << 28
//   cold[cap] (16 B)  x, y, z = outcome_probs[0..3] (f32 sums L, D, W), w = parent
// The children of a node are contiguous in BOTH arrays, so one level of selection reads
// <= 144 contiguous bytes (4.5 sectors instead of 9).  The cold half is touched only by backprop
// (and target_q): it is written the first time a node is backed up and never at expansion.
// y is recomputed with the same f32 subtraction, from the same operands, that q() would perform
// at selection time, so every comparison sees the bits it saw before.
#pragma once
#include "tpg2.cuh"

namespace tp3 {

using namespace eng;
using tp2::Pend;
using tp2::RoundCnt;
using tp2::RootOut;
using tp2::K_INIT;
using tp2::K_LEAF;
using tp2::K_NONE;
using tp2::K_TERMINAL;

constexpr uint32_t M24 = 0x00ffffffu;

__device__ __forceinline__ uint4 ld16(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t h_vis(const uint4& h) { return h.x & M24; }
__device__ __forceinline__ uint32_t h_sol(const uint4& h) { return h.x >> 24; }
__device__ __forceinline__ uint32_t h_fc(const uint4& h) { return h.w & M24; }
__device__ __forceinline__ uint32_t h_nch(const uint4& h) { return (h.w >> 24) & 15u; }
__device__ __forceinline__ uint32_t h_act(const uint4& h) { return h.w >> 28; }

struct Arena {
    uint4* hot;
    uint4* cold;
};

// One explore from the (already loaded) root up to the point where the policy is needed
// (mcts.rs:310-325, 327-372, 374-406).  pd.lc carries the parent of pd.id in its upper bits for
// K_TERMINAL (see backprop); K_LEAF keeps it in `leaf_parent`.
template <int CW, int FPU>
__device__ __forceinline__ int descend(const KParams& p, uint32_t* ss, const Arena& a, uint32_t& nn, const uint4& root, uint64_t& my, uint64_t& op,
                                       Pend& pd, uint32_t& parent_of_id, RoundCnt& rc) {
    const syn_mcts_cfg& cfg = p.cfg.mcts;
    constexpr bool PQ = FPU == SYN_FPU_PARENT_Q;
    uint32_t cur = 0u, par = 0u;
    uint4 c = root; // hot record of the current node
    uint32_t depth = 0;
    const bool puct = cfg.exploration_kind == SYN_EXPLORATION_POLYNOMIAL_UCT;
    for (;;) {
        const uint32_t sol = h_sol(c), nch = h_nch(c);
        if (sol) { rc.levels = depth; pd.kind = K_TERMINAL; pd.id = cur; pd.fc = sol; parent_of_id = par; return 0; } // mcts.rs:314-316
        if (nch == 0u) break;
        // ---- select_best_child (mcts.rs:327-372): first strict maximum in child order
        const float cvis = (float)h_vis(c);
        const float pterm = puct ? __fsqrt_rn(cvis) : __fsqrt_rn(__fmul_rn(cfg.c, syn_logf(cvis)));
        const float parent_q = PQ ? __fdiv_rn(__uint_as_float(c.y), cvis) : 0.0f; // Fpu::ParentQ = parent.q() (mcts.rs:353)
        const uint32_t cfc = h_fc(c);
        uint32_t b = 0u;
        float bval = 0.0f;
        uint4 bh = make_uint4(0u, 0u, 0u, 0u);
        for (uint32_t k0 = 0; k0 < nch; k0 += (uint32_t)CW) { // CW hot records per memory round trip
            uint4 chs[CW];
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)CW; ++j) chs[j] = ld16(a.hot + cfc + min(k0 + j, nch - 1u));
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)CW; ++j) {
                const uint32_t k = k0 + j;
                const uint4& ch = chs[j];
                const uint32_t csol = h_sol(ch), cn = h_nch(ch);
                const float vis = (float)h_vis(ch);
                float q;
                if (csol) {
                    uint32_t kd = sol_kind(csol);
                    q = cfg.select_solved_nodes ? (kd == SYN_KIND_WIN ? -1.0f : (kd == SYN_KIND_LOSE ? 1.0f : 0.0f)) : __uint_as_float(0xff800000u);
                } else if (cn == 0u) {
                    if (PQ) q = parent_q;
                    else if (FPU == SYN_FPU_CONST) q = cfg.fpu_a;
                    else q = k < nch ? tp2::fpu_normal_draw(p, ss) : 0.0f;
                } else {
                    q = -__fdiv_rn(__uint_as_float(ch.y), vis);
                }
                float u = puct ? __fdiv_rn(__fmul_rn(__fmul_rn(cfg.c, __uint_as_float(ch.z)), pterm), __fadd_rn(1.0f, vis))
                               : __fdiv_rn(pterm, __fsqrt_rn(vis));
                float value = __fadd_rn(q, u);
                if (k < nch && (k == 0u || value > bval)) { b = k; bval = value; bh = ch; }
            }
        }
        rc.scanned += nch;
        par = cur;
        cur = cfc + b;
        c = bh;
        c4::step(my, op, (int)h_act(c));
        if (++depth >= 64u) return DERR_DEPTH_OVERFLOW;
    }
    rc.levels = depth; // one select_best_child call per level walked
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
        uint32_t fc = nn;
        if (fc + n > p.arena_nodes || fc + n > M24) return DERR_ARENA_OVERFLOW;
        reinterpret_cast<uint32_t*>(a.hot + cur)[3] = fc | (n << 24) | (c.w & 0xf0000000u); // mark_visited (mcts.rs:399-400)
        nn = fc + n;
        rc.expansions += 1u;
        rc.created += n;
        if (cfg.auto_extend && n == 1u) { // mcts.rs:404-405: recurse into the only child, no policy call
            int only = __ffs(lm) - 1;
            uint32_t s2 = (cs2 >> (2 * only)) & 3u;
            uint32_t osol = s2 == 1u ? c4::SOL_LOSE0 : (s2 == 2u ? c4::SOL_DRAW0 : 0u);
            // `cur` becomes an inner node of this explore's path without ever having been backed up:
            // give its cold half the parent link backprop will follow (sums start at zero)
            if (h_vis(c) == 0u) a.cold[cur] = make_uint4(0u, 0u, 0u, par);
            c = make_uint4(osol << 24, 0u, __float_as_uint(1.0f), (uint32_t)only << 28);
            a.hot[fc] = c;
            c4::step(my, op, only);
            par = cur;
            cur = fc;
            if (++depth >= 64u) return DERR_DEPTH_OVERFLOW;
            if (osol) { pd.kind = K_TERMINAL; pd.id = cur; pd.fc = osol; parent_of_id = par; return 0; } // mcts.rs:377-379
            continue;
        }
        pd.kind = K_LEAF; pd.id = cur; pd.fc = fc; pd.lc = lm | (cs2 << 9);
        parent_of_id = par;
        return 0;
    }
}

// mcts.rs:429-488 from node `id` up to the root.  `parent_of_id` is used when `id` has never been
// backed up (its cold half does not exist yet).
__device__ __forceinline__ uint32_t backprop(const syn_mcts_cfg& cfg, const Arena& a, uint32_t id, uint32_t parent_of_id, float v0, float v1, float v2,
                                             bool solved) {
    uint32_t levels = 0;
    bool first = true;
    for (;;) {
        uint4 h = ld16(a.hot + id);
        uint4 cd = ld16(a.cold + id); // requested together with the hot half; ignored for a never-visited first node
        ++levels;
        const uint32_t visi = h_vis(h);
        float o0 = __uint_as_float(cd.x), o1 = __uint_as_float(cd.y), o2 = __uint_as_float(cd.z);
        uint32_t par = cd.w;
        if (first && visi == 0u) { o0 = o1 = o2 = 0.0f; par = parent_of_id; }
        first = false;
        const float vis = (float)visi;
        uint32_t nsol = h_sol(h);
        if (cfg.solve && solved) {
            const uint32_t nch = h_nch(h), fc = h_fc(h);
            uint32_t bk = sol_key(nsol);
            bool all_solved = true;
            for (uint32_t k = 0; k < nch; ++k) {
                uint32_t csol = reinterpret_cast<const uint32_t*>(a.hot + fc + k)[0] >> 24;
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
                    v0 = -o0; v1 = -o1; v2 = -o2;
                    float add = vis + 1.0f;
                    if (slot == 2) v2 = v2 + add;
                    else if (slot == 1) v1 = v1 + add;
                    else v0 = v0 + add;
                }
                nsol = best;
            } else {
                solved = false;
            }
        }
        const float n0 = o0 + v0, n1 = o1 + v1, n2 = o2 + v2;
        a.cold[id] = make_uint4(__float_as_uint(n0), __float_as_uint(n1), __float_as_uint(n2), par);
        *reinterpret_cast<uint2*>(a.hot + id) = make_uint2((visi + 1u) | (nsol << 24), __float_as_uint(__fsub_rn(n2, n0)));
        if (id == 0u) break;
        float tmp = v0; v0 = v2; v2 = tmp;
        id = par;
    }
    return levels;
}

// The rest of visit() after Policy::eval (mcts.rs:384-397 child records, 409-423 stable softmax over
// the legal children in child order).  Only the hot halves are written.
__device__ __forceinline__ void write_children(const Arena& a, const Pend& pd, const float (&logits)[9]) {
    const uint32_t legal = pd.lc & 0x1ffu, csol2 = pd.lc >> 9;
    float e[9];
    float total = 0.0f;
    float mx = __uint_as_float(0xff800000u);
#pragma unroll
    for (int col = 0; col < 9; ++col)
        if ((legal >> col) & 1u) mx = fmaxf(mx, logits[col]);
#pragma unroll
    for (int col = 0; col < 9; ++col) {
        e[col] = 0.0f;
        if ((legal >> col) & 1u) {
            e[col] = syn_expf(__fsub_rn(logits[col], mx));
            total = __fadd_rn(total, e[col]);
        }
    }
    uint32_t rank = 0u;
#pragma unroll
    for (int col = 0; col < 9; ++col) {
        if ((legal >> col) & 1u) {
            uint32_t s2 = (csol2 >> (2 * col)) & 3u;
            uint32_t csol = s2 == 1u ? c4::SOL_LOSE0 : (s2 == 2u ? c4::SOL_DRAW0 : 0u);
            a.hot[pd.fc + rank] = make_uint4(csol << 24, 0u, __float_as_uint(__fdiv_rn(e[col], total)), (uint32_t)col << 28);
            ++rank;
        }
    }
}

// mcts.rs:229-269 after the construction visit.  Cold path.
__device__ __noinline__ void add_root_noise(const KParams& p, uint32_t* ss, uint4* hot) {
    const syn_mcts_cfg& cfg = p.cfg.mcts;
    if (cfg.noise_kind == SYN_NOISE_NONE) return;
    uint4 r = ld16(hot);
    uint32_t nch = h_nch(r);
    if (nch < 2u) return;
    float w = cfg.noise_weight;
    float vals[9];
    for (int k = 0; k < 9; ++k) vals[k] = __fdiv_rn(1.0f, (float)nch);
    if (cfg.noise_kind == SYN_NOISE_DIRICHLET) {
        rng::Stream st;
        st.init(tp2::stream_seed(p, ss[tp2::SS_GI], SYN_STREAM_NOISE), ss[tp2::SS_NOISE_POS]);
        syn_dirichlet(st, cfg.noise_alpha, (int)nch, vals);
        ss[tp2::SS_NOISE_POS] = (uint32_t)st.pos;
    }
    for (uint32_t k = 0; k < nch; ++k) {
        uint32_t* hw = reinterpret_cast<uint32_t*>(hot + h_fc(r) + k);
        float pr = __uint_as_float(hw[2]);
        pr = __fadd_rn(__fmul_rn(pr, __fsub_rn(1.0f, w)), __fmul_rn(w, vals[k]));
        hw[2] = __float_as_uint(pr);
    }
}

// What the driver reads from a finished tree (mcts.rs:174-225, 273-306), by COLUMN.
struct ReadRoot3 {
    __device__ __forceinline__ void operator()(const uint4* hot, uint32_t cap, uint32_t action_selection, RootOut& r) const {
        const uint4* cold = hot + cap;
        uint4 root = ld16(hot), rc = ld16(cold);
        uint32_t nch = h_nch(root), rsol = h_sol(root);
        const float rvis = (float)h_vis(root);
        for (int k = 0; k < 9; ++k) { r.pi[k] = 0.0f; r.visits[k] = 0.0f; r.child_sol[k] = 0u; }
        float total = 0.0f, b0 = 0.0f, b1 = 0.0f;
        int best = 0;
        uint32_t legal = 0u;
        for (uint32_t k = 0; k < nch; ++k) {
            uint4 ch = ld16(hot + h_fc(root) + k);
            uint32_t csol = h_sol(ch), act = h_act(ch);
            const float cvis = (float)h_vis(ch);
            float v; // target_policy (mcts.rs:174-211)
            if (rvis == 1.0f) v = sol_kind(rsol) == SYN_KIND_WIN ? (sol_kind(csol) == SYN_KIND_LOSE ? 1.0f : 0.0f) : 1.0f;
            else v = cvis;
            total = __fadd_rn(total, v);
            float k0, k1; // best_action (mcts.rs:273-294): key (k0, k1), strict lexicographic >, first child incumbent
            uint32_t kind = sol_kind(csol);
            if (kind == SYN_KIND_WIN) { k0 = 0.0f; k1 = (float)(csol & 63u); }
            else if (kind == 0u) { k0 = 1.0f; k1 = action_selection == SYN_ACTION_Q ? -__fdiv_rn(__uint_as_float(ch.y), cvis) : cvis; }
            else if (kind == SYN_KIND_DRAW) { k0 = 2.0f; k1 = -(float)(csol & 63u); }
            else { k0 = 3.0f; k1 = -(float)(csol & 63u); }
            if (k == 0u || k0 > b0 || (k0 == b0 && k1 > b1)) { b0 = k0; b1 = k1; best = (int)act; }
            legal |= 1u << act;
#pragma unroll
            for (int col = 0; col < 9; ++col)
                if ((int)act == col) { r.pi[col] = v; r.visits[col] = cvis; r.child_sol[col] = csol; }
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
            r.q0 = __fdiv_rn(__uint_as_float(rc.x), rvis); r.q1 = __fdiv_rn(__uint_as_float(rc.y), rvis); r.q2 = __fdiv_rn(__uint_as_float(rc.z), rvis);
        }
    }
};

} // namespace tp3

namespace eng {

// One persistent CTA per SM, TEAMS teams of 128 threads sharing SLOTS MLP slots (mlp_team.cuh).
template <int TEAMS, int SLOTS, int CW, bool PROF>
__global__ void __launch_bounds__(128 * TEAMS, 1) selfplay_nn_tpg3_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ unsigned long long s_cnt[CNT_ALL];
    mlpteam::Smem<TEAMS, SLOTS>& ms = *reinterpret_cast<mlpteam::Smem<TEAMS, SLOTS>*>(smem_raw);
    if (threadIdx.x < CNT_ALL) s_cnt[threadIdx.x] = 0ull;
    mlpteam::setup<TEAMS, SLOTS>(ms, p.weight_image);
    const int team = threadIdx.x >> 7, r = threadIdx.x & 127;
    const size_t slot_id = (size_t)blockIdx.x * (128 * TEAMS) + threadIdx.x;
    uint32_t* const ss = p.slot_state + tp2::SS_WORDS * slot_id;
    const syn_mcts_cfg& cfg = p.cfg.mcts;
    const uint32_t stop_vis = p.cfg.num_explores + 1u; // explore_n is over when the root has 1 + num_explores visits
    tp3::Arena a;
    a.hot = p.nodes + 2 * slot_id * p.arena_nodes;
    a.cold = a.hot + p.arena_nodes;
    uint32_t nn = 1u;
    int phase = PH_NEED_GAME;
    long long t_adv = 0, t_wait = 0, t_mlp = 0, t_fin = 0, t_start = PROF ? clock64() : 0;
    uint32_t rounds = 0, leaves = 0;
    for (;;) {
        long long t0 = PROF ? clock64() : 0;
        // ---- cold bookkeeping, then at most one descent
        tp2::Pend pd;
        pd.kind = tp2::K_NONE; pd.id = 0u; pd.fc = 0u; pd.lc = 0u;
        uint32_t parent_of_id = 0u;
        tp2::RoundCnt rc = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
        uint64_t my = 0, op = 0;
        if (phase != PH_DONE) {
            int err = 0;
            uint4 root = make_uint4(0u, 0u, 0u, 0u);
            bool go = true;
            if (phase == PH_EXPLORE) { // explore_n (mcts.rs:139-147): stop at num_explores or once the root is solved
                root = tp3::ld16(a.hot);
                my = tp2::ss_load64(ss, tp2::SS_MY); op = tp2::ss_load64(ss, tp2::SS_OP);
                if (tp3::h_vis(root) >= stop_vis || tp3::h_sol(root) != 0u) {
                    int pe = tp2::end_of_move(p, ss, a.hot, nn, tp3::h_vis(root) - 1u, tp3::ReadRoot3());
                    phase = pe & 0xff; err = pe >> 8;
                    go = false; // the next tree starts next round
                }
            } else {
                if (phase == PH_NEED_GAME) phase = tp2::next_game(p, ss);
                if (phase == PH_NEW_TREE) { // MCTS::with_capacity (mcts.rs:123-137): fresh arena, root only
                    a.hot[0] = root;
                    my = tp2::ss_load64(ss, tp2::SS_MY); op = tp2::ss_load64(ss, tp2::SS_OP);
                    nn = 1u;
                    atomicAdd(&s_cnt[CNT_TREES], 1ull);
                    phase = PH_EXPLORE;
                } else {
                    go = false; // PH_DONE
                }
            }
            if (go && !err) {
                const uint32_t init = tp3::h_vis(root) == 0u ? (uint32_t)tp2::K_INIT : 0u; // the construction visit (mcts.rs:133)
                if (cfg.fpu_kind == SYN_FPU_CONST) err = tp3::descend<CW, SYN_FPU_CONST>(p, ss, a, nn, root, my, op, pd, parent_of_id, rc);
                else if (cfg.fpu_kind == SYN_FPU_PARENT_Q) err = tp3::descend<CW, SYN_FPU_PARENT_Q>(p, ss, a, nn, root, my, op, pd, parent_of_id, rc);
                else err = tp3::descend<CW, SYN_FPU_NORMAL>(p, ss, a, nn, root, my, op, pd, parent_of_id, rc);
                pd.kind |= init;
            }
            if (err) { atomicCAS(p.error, 0, err); phase = PH_DONE; pd.kind = tp2::K_NONE; }
        }
        __syncwarp();
        { // statistics of the descent (summed per warp here so that they are not live across the forward)
            uint32_t a0 = __reduce_add_sync(0xffffffffu, rc.levels), a1 = __reduce_add_sync(0xffffffffu, rc.scanned);
            uint32_t a2 = __reduce_add_sync(0xffffffffu, rc.expansions), a3 = __reduce_add_sync(0xffffffffu, rc.created);
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&s_cnt[CNT_SELECT_LEVELS], (unsigned long long)a0); atomicAdd(&s_cnt[CNT_CHILDREN_SCANNED], (unsigned long long)a1);
                atomicAdd(&s_cnt[CNT_EXPANSIONS], (unsigned long long)a2); atomicAdd(&s_cnt[CNT_CHILDREN_CREATED], (unsigned long long)a3);
            }
        }
        long long t1 = PROF ? clock64() : 0;
        const bool need = (pd.kind & tp2::K_LEAF) != 0u;
        if (PROF) leaves += (uint32_t)__popc(__ballot_sync(0xffffffffu, need));
        if (!mlpteam::team_any(team, phase != PH_DONE)) break; // no thread of this team has a game left
        uint32_t mma_phase;
        const int slot = mlpteam::acquire_slot<TEAMS, SLOTS>(ms, team, r, mma_phase);
        long long t2 = PROF ? clock64() : 0;
        if (need) mlpteam::write_features(ms.a[slot], ms.col_lut, r, my, op);
        float y[12];
        mlpteam::forward<TEAMS, SLOTS>(ms, team, slot, r, mma_phase, y);
        mlpteam::release_slot<TEAMS, SLOTS>(ms, team, r, slot, mma_phase);
        long long t3 = PROF ? clock64() : 0;
        // ---- finish: child records for leaves, then ONE backprop site for every kind of explore
        uint32_t bp_levels = 0u;
        if (pd.kind != tp2::K_NONE) {
            float v0, v1, v2;
            bool solved;
            if (need) {
                // value.softmax(-1) (study-connect4/src/policies.rs:54-56)
                float m = fmaxf(y[9], fmaxf(y[10], y[11]));
                float e0 = syn_expf(__fsub_rn(y[9], m)), e1 = syn_expf(__fsub_rn(y[10], m)), e2 = syn_expf(__fsub_rn(y[11], m));
                float tot = __fadd_rn(__fadd_rn(e0, e1), e2);
                float lg[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) lg[k] = y[k];
                tp3::write_children(a, pd, lg);
                v0 = __fdiv_rn(e0, tot); v1 = __fdiv_rn(e1, tot); v2 = __fdiv_rn(e2, tot);
                solved = (pd.lc >> 9) != 0u;
            } else {
                int idx = sol_index(pd.fc);
                v0 = idx == 0 ? 1.0f : 0.0f; v1 = idx == 1 ? 1.0f : 0.0f; v2 = idx == 2 ? 1.0f : 0.0f;
                solved = true;
            }
            bp_levels = tp3::backprop(cfg, a, pd.id, parent_of_id, v0, v1, v2, solved);
            if (pd.kind & tp2::K_INIT) tp3::add_root_noise(p, ss, a.hot);
        }
        __syncwarp();
        { // statistics of the finish: one shared-memory atomic per warp and counter
            uint32_t a4 = __reduce_add_sync(0xffffffffu, bp_levels), a5 = __reduce_add_sync(0xffffffffu, need ? 1u : 0u);
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&s_cnt[CNT_BACKPROP_LEVELS], (unsigned long long)a4); atomicAdd(&s_cnt[CNT_LEAF_EVALS], (unsigned long long)a5);
            }
        }
        if (PROF) {
            long long t4 = clock64();
            t_adv += t1 - t0; t_wait += t2 - t1; t_mlp += t3 - t2; t_fin += t4 - t3; ++rounds;
        }
    }
    if (PROF && (threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt[DBG_T_ADVANCE], (unsigned long long)t_adv);
        atomicAdd(&s_cnt[DBG_T_TEAMWAIT], (unsigned long long)t_wait);
        atomicAdd(&s_cnt[DBG_T_MLP], (unsigned long long)t_mlp);
        atomicAdd(&s_cnt[DBG_T_FINISH], (unsigned long long)t_fin);
        atomicAdd(&s_cnt[DBG_ROUNDS], (unsigned long long)rounds);
        atomicAdd(&s_cnt[DBG_LEAVES], (unsigned long long)leaves);
        atomicAdd(&s_cnt[DBG_T_TOTAL], (unsigned long long)(clock64() - t_start));
    }
    mlpteam::teardown<TEAMS, SLOTS>(ms); // ends with a CTA barrier: every warp's counters are in s_cnt
    __syncthreads();
    if (threadIdx.x < CNT_ALL && s_cnt[threadIdx.x]) atomicAdd(p.counters + threadIdx.x, s_cnt[threadIdx.x]);
}

} // namespace eng
