// tpg4.cuh — the kernels of the tpg4 schedule: tpg4_tree.cuh's family blocks (select_best_child reads ONE 128-byte
// line per level, -q memoised, the walk up's lines prefetched during the leaf evaluation) under tpg2.cuh's round structure,
// with tpg2.cuh's seating.  See tpg4_tree.cuh for why (the L2-miss sector rate of the memory system is the wall).
//
// Replaces synthesis/src/mcts.rs:29-489 and synthesis/src/alpha_zero.rs:229-338 of the reference.
#pragma once
#include "tpg2_split.cuh"
#include "tpg4_tree.cuh"

namespace tp4 {

#ifndef SYN_BP4_CHUNK
#define SYN_BP4_CHUNK 8
#endif
constexpr int BP_CHUNK = SYN_BP4_CHUNK; // path nodes the walk up loads together

// mcts.rs:229-269 after the construction visit.  Cold path.
__device__ __noinline__ void add_root_noise(const KParams& p, uint32_t* ss, uint4* arena) {
    const syn_mcts_cfg& cfg = p.cfg.mcts;
    if (cfg.noise_kind == SYN_NOISE_NONE) return;
    const uint32_t rlink = line0(arena, 0u)[link_word(ROOT_ID)];
    const uint32_t fc = link_fc(rlink), nch = link_nch(rlink);
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
    uint32_t* l0 = line0(arena, fc);
    uint32_t k = 0u; // the k-th child in child order = the k-th column with a child
    for (uint32_t col = 0; col < 9u; ++col) {
        if (l0[W_LINK + col] == ABSENT) continue;
        float pr = __uint_as_float(l0[W_PRIOR + col]);
        pr = __fadd_rn(__fmul_rn(pr, __fsub_rn(1.0f, w)), __fmul_rn(w, vals[k]));
        l0[W_PRIOR + col] = __float_as_uint(pr);
        ++k;
    }
}

struct ReadRoot4 { // end_of_move's reader: the root position tells which column each child of the root plays
    uint64_t occ;
    __device__ __forceinline__ void operator()(const uint4* arena, uint32_t cap, uint32_t action_selection, RootOut& r) const {
        tp4::read_root(arena, occ, action_selection, r);
    }
};

} // namespace tp4

namespace eng {

// One persistent CTA per SM, TEAMS teams of 128 threads sharing SLOTS MLP slots (mlp_team.cuh); tpg4_tree.cuh's family
// blocks.  A round is [cold bookkeeping] -> descend (+ prefetch of what the walk up will read) -> team forward on
// tcgen05 -> finish (children block, the walk up).
// FPU = the configured syn_fpu_kind: a kernel per kind, so that the common Fpu::Const build carries neither ParentQ's
// state nor Normal's call in the child loop.
template <int TEAMS, bool PROF, int FPU>
__global__ void __launch_bounds__(128 * TEAMS, 1) selfplay_nn_tpg4_kernel(const __grid_constant__ KParams p) {
    constexpr int SLOTS = TEAMS < 2 ? TEAMS : 2; // mlp_split.cuh: the fp32-grade forward, activations in tensor memory
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ unsigned long long s_cnt[CNT_ALL];
    __shared__ unsigned long long s_wcnt[4 * TEAMS][tp2::WC_N]; // per-warp statistics rows (plain adds by lane 0)
    mlps::Smem<TEAMS, SLOTS>& ms = *reinterpret_cast<mlps::Smem<TEAMS, SLOTS>*>(smem_raw);
    constexpr int NT = 128 * TEAMS, PATH_CAP = tp2s::PATH_CAP;
    uint32_t* const path = reinterpret_cast<uint32_t*>(smem_raw + sizeof(mlps::Smem<TEAMS, SLOTS>)) + threadIdx.x; // [PATH_CAP][NT] after the MLP state
    if (threadIdx.x < CNT_ALL) s_cnt[threadIdx.x] = 0ull;
    for (int i = threadIdx.x; i < 4 * TEAMS * tp2::WC_N; i += 128 * TEAMS) (&s_wcnt[0][0])[i] = 0ull;
    unsigned long long* const wc = s_wcnt[threadIdx.x >> 5];
    mlps::setup<TEAMS, SLOTS>(ms, p.weight_image, p.weight_image_lo);
    const int team = threadIdx.x >> 7, r = threadIdx.x & 127;
    const tp2::Seat seat = tp2::seat_of(p, team, r, 4);
    uint32_t* const ss = p.slot_state + tp2::SS_WORDS * seat.slot;
    const syn_mcts_cfg& cfg = p.cfg.mcts;
    const float stop_vis = (float)(p.cfg.num_explores + 1u); // explore_n is over when the root has 1 + num_explores visits
    uint4* const nodes = p.nodes + 2 * seat.slot * p.arena_nodes; // arena_nodes * 32 bytes = arena_nodes / 4 lines of 128 bytes
    const uint32_t arena_lines = p.arena_nodes >> 2;
    uint32_t nn = 1u, nb = 1u;
    int phase = seat.active ? PH_NEED_GAME : PH_DONE;
    // per-warp phase clocks (syn_engine_debug_counters): only in the PROF instantiation
    long long t_adv = 0, t_wait = 0, t_mlp = 0, t_fin = 0, t_start = PROF ? clock64() : 0;
    uint32_t rounds = 0, leaves = 0;
    for (;;) {
        long long t0 = PROF ? clock64() : 0;
        // ---- cold bookkeeping, then at most one descent
        tp2::Pend pd;
        pd.kind = tp2::K_NONE; pd.id = 0u; pd.fc = 0u; pd.lc = 0u; pd.depth = 0u;
        tp2::RoundCnt rc = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
        uint64_t my = 0, op = 0;
        if (phase != PH_DONE) {
            int err = 0;
            tp4::Cur root;
            bool go = true;
            if (phase == PH_EXPLORE) { // explore_n (mcts.rs:139-147): stop at num_explores or once the root is solved
                root = tp4::load_root(nodes);
                my = tp2::ss_load64(ss, tp2::SS_MY); op = tp2::ss_load64(ss, tp2::SS_OP);
                if (root.vis >= stop_vis || tp4::link_sol(root.link) != 0u) {
                    int pe = tp2::end_of_move(p, ss, nodes, nn, (uint32_t)root.vis - 1u, tp4::ReadRoot4{my | op});
                    phase = pe & 0xff; err = pe >> 8;
                    go = false; // the next tree starts next round
                }
            } else {
                if (phase == PH_NEED_GAME) phase = tp2::next_game(p, ss);
                if (phase == PH_NEW_TREE) { // MCTS::with_capacity (mcts.rs:123-137): fresh arena, root only
                    root.id = tp4::ROOT_ID; root.vis = root.nq = 0.0f; root.link = 0u;
                    tp4::store_empty_root(nodes);
                    my = tp2::ss_load64(ss, tp2::SS_MY); op = tp2::ss_load64(ss, tp2::SS_OP);
                    nn = 1u; nb = 1u;
                    atomicAdd(&s_cnt[CNT_TREES], 1ull);
                    phase = PH_EXPLORE;
                } else {
                    go = false; // PH_DONE
                }
            }
            if (go && !err) {
                const uint32_t init = root.vis == 0.0f ? (uint32_t)tp2::K_INIT : 0u; // the construction visit (mcts.rs:133)
                err = tp4::descend<FPU, NT, PATH_CAP>(cfg, arena_lines, tp2::FpuDraw{p, ss}, nodes, nb, nn, root, my, op, pd, rc, path);
                if (!err) tp4::prefetch_path<NT, PATH_CAP>(nodes, path, pd); // the walk up comes after the leaf evaluation: its lines can travel meanwhile
                pd.kind |= init;
            }
            if (err) { atomicCAS(p.error, 0, err); phase = PH_DONE; pd.kind = tp2::K_NONE; }
        }
        __syncwarp();
        { // statistics of the descent (summed per warp here so that they are not live across the forward)
            uint32_t a0 = __reduce_add_sync(0xffffffffu, rc.levels), a1 = __reduce_add_sync(0xffffffffu, rc.scanned);
            uint32_t a2 = __reduce_add_sync(0xffffffffu, rc.expansions), a3 = __reduce_add_sync(0xffffffffu, rc.created);
            if ((threadIdx.x & 31) == 0) { wc[tp2::WC_LEVELS] += a0; wc[tp2::WC_SCANNED] += a1; wc[tp2::WC_EXPANSIONS] += a2; wc[tp2::WC_CREATED] += a3; }
        }
        long long t1 = PROF ? clock64() : 0;
        const bool need = (pd.kind & tp2::K_LEAF) != 0u;
        if (PROF) leaves += (uint32_t)__popc(__ballot_sync(0xffffffffu, need));
        if (!mlps::team_any(team, phase != PH_DONE)) break; // no thread of this team has a game left
        uint32_t mma_phase;
        const int slot = mlps::acquire_slot<TEAMS, SLOTS>(ms, team, r, mma_phase);
        long long t2 = PROF ? clock64() : 0;
        mlps::write_features<TEAMS, SLOTS>(ms, slot, r, my, op, need);
        float y[12];
        mlps::forward<TEAMS, SLOTS>(ms, p.mlp_bias, team, slot, r, mma_phase, y);
        mlps::release_slot<TEAMS, SLOTS>(ms, team, r, slot, mma_phase);
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
                tp4::write_children(nodes, pd, lg);
                v0 = __fdiv_rn(e0, tot); v1 = __fdiv_rn(e1, tot); v2 = __fdiv_rn(e2, tot);
                solved = (pd.lc >> 9) != 0u;
            } else {
                int idx = sol_index(pd.fc);
                v0 = idx == 0 ? 1.0f : 0.0f; v1 = idx == 1 ? 1.0f : 0.0f; v2 = idx == 2 ? 1.0f : 0.0f;
                solved = true;
            }
            tp4::backprop<NT, PATH_CAP, tp4::BP_CHUNK>(cfg, nodes, path, pd.depth, pd.id, v0, v1, v2, solved);
            bp_levels = pd.depth + 1u; // mcts.rs:431: every level from the leaf to the root
            if (pd.kind & tp2::K_INIT) tp4::add_root_noise(p, ss, nodes);
        }
        __syncwarp();
        { // statistics of the finish
            uint32_t a4 = __reduce_add_sync(0xffffffffu, bp_levels), a5 = __reduce_add_sync(0xffffffffu, need ? 1u : 0u);
            if ((threadIdx.x & 31) == 0) { wc[tp2::WC_BACKPROP] += a4; wc[tp2::WC_LEAF_EVALS] += a5; }
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
    mlps::teardown<TEAMS, SLOTS>(ms); // ends with a CTA barrier: every warp's counters are in s_cnt / s_wcnt
    __syncthreads();
    if (threadIdx.x < CNT_ALL && s_cnt[threadIdx.x]) atomicAdd(p.counters + threadIdx.x, s_cnt[threadIdx.x]);
    if (threadIdx.x < tp2::WC_N) {
        unsigned long long t = 0ull;
        for (int w = 0; w < 4 * TEAMS; ++w) t += s_wcnt[w][threadIdx.x];
        if (t) atomicAdd(p.counters + tp2::wc_counter(threadIdx.x), t);
    }
}

} // namespace eng
