// tpg2_split.cuh — selfplay_nn_tpg2_kernel with the fp32-grade Connect4Net forward of mlp_split.cuh (split-fp16 operands,
// activations resident in tensor memory): the product kernel for network leaves.  The tree code, the round structure and the
// seating are tpg2.cuh's, line for line; only the leaf evaluation differs.
//
// Replaces synthesis/src/mcts.rs:29-489, synthesis/src/alpha_zero.rs:229-338 and study-connect4/src/policies.rs:28-59.
#pragma once
#include "mlp_split.cuh"
#include "tpg2.cuh"

namespace tp2s {
constexpr int PATH_CAP = 12; // levels of the path table per thread: shared memory no longer holds activation tiles
template <int TEAMS>
constexpr size_t smem_bytes() { return sizeof(mlps::Smem<TEAMS, (TEAMS < 2 ? TEAMS : 2)>) + (size_t)PATH_CAP * 128 * TEAMS * sizeof(uint32_t); }
} // namespace tp2s

namespace eng {

// One persistent CTA per SM, TEAMS teams of 128 threads sharing two MLP slots of tensor memory (mlp_split.cuh).
template <int TEAMS, bool PROF, int FPUK = -1> // FPUK: see selfplay_nn_tpg2_kernel
__global__ void __launch_bounds__(128 * TEAMS, 1) selfplay_nn_tpg2s_kernel(const __grid_constant__ KParams p) {
    constexpr int SLOTS = TEAMS < 2 ? TEAMS : 2;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ unsigned long long s_cnt[CNT_ALL];
    // the six per-round statistics are kept per WARP (plain 64-bit adds by lane 0, nobody else touches the row): shared-memory
    // 64-bit atomics on six CTA-wide counters were 3.8 % of the stall samples (profiles/r1s_*)
    __shared__ unsigned long long s_wcnt[4 * TEAMS][tp2::WC_N];
    mlps::Smem<TEAMS, SLOTS>& ms = *reinterpret_cast<mlps::Smem<TEAMS, SLOTS>*>(smem_raw);
    constexpr int NT = 128 * TEAMS, PATH_CAP = tp2s::PATH_CAP;
    uint32_t* const path = reinterpret_cast<uint32_t*>(smem_raw + sizeof(mlps::Smem<TEAMS, SLOTS>)) + threadIdx.x; // [PATH_CAP][NT] after the MLP state
    if (threadIdx.x < CNT_ALL) s_cnt[threadIdx.x] = 0ull;
    for (int i = threadIdx.x; i < 4 * TEAMS * tp2::WC_N; i += 128 * TEAMS) (&s_wcnt[0][0])[i] = 0ull;
    unsigned long long* const wc = s_wcnt[threadIdx.x >> 5];
    mlps::setup<TEAMS, SLOTS>(ms, p.weight_image, p.weight_image_lo);
    const int team = threadIdx.x >> 7, r = threadIdx.x & 127;
    const tp2::Seat seat = tp2::seat_of(p, team, r, 4);
    const size_t slot_id = seat.slot;
    uint32_t* const ss = p.slot_state + tp2::SS_WORDS * slot_id;
    const syn_mcts_cfg& cfg = p.cfg.mcts;
    const float stop_vis = (float)(p.cfg.num_explores + 1u); // explore_n is over when the root has 1 + num_explores visits
    constexpr int CW = TEAMS >= 5 ? 3 : 5;
    tp2::Game g;
    g.nodes = p.nodes + 2 * slot_id * p.arena_nodes;
    g.nn = 1u; g.phase = seat.active ? PH_NEED_GAME : PH_DONE; g.slow = p.no_reductions != 0u;
    // per-warp phase clocks (syn_engine_debug_counters): only in the PROF instantiation, they cost 14 registers
    long long t_adv = 0, t_wait = 0, t_mlp = 0, t_fin = 0, t_start = PROF ? clock64() : 0;
    uint32_t rounds = 0, leaves = 0;
    for (;;) {
        long long t0 = PROF ? clock64() : 0;
        // ---- cold bookkeeping, then at most one descent
        tp2::Pend pd;
        pd.kind = tp2::K_NONE; pd.id = 0u; pd.fc = 0u; pd.lc = 0u; pd.depth = 0u;
        tp2::RoundCnt rc = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
        uint64_t my = 0, op = 0;
        if (g.phase != PH_DONE) {
            int err = 0;
            tp2::Rec root;
            bool go = true;
            if (g.phase == PH_EXPLORE) { // explore_n (mcts.rs:139-147): stop at num_explores or once the root is solved
                root = tp2::load_rec(g.nodes, 0u);
                my = tp2::ss_load64(ss, tp2::SS_MY); op = tp2::ss_load64(ss, tp2::SS_OP);
                if (root.vis >= stop_vis || ((root.pk >> 8) & 0xffu) != 0u) {
                    int pe = tp2::end_of_move(p, ss, g.nodes, g.nn, (uint32_t)root.vis - 1u, tp2::ReadRoot2());
                    g.phase = pe & 0xff; err = pe >> 8;
                    go = false; // the next tree starts next round
                }
            } else {
                if (g.phase == PH_NEED_GAME) g.phase = tp2::next_game(p, ss);
                if (g.phase == PH_NEW_TREE) { // MCTS::with_capacity (mcts.rs:123-137): fresh arena, root only
                    root.vis = root.o0 = root.o1 = root.o2 = 0.0f;
                    root.prior = root.parent = root.fc = root.pk = 0u;
                    tp2::store_rec(g.nodes, 0u, 0.f, 0.f, 0.f, 0.f, 0u, 0u, 0u, 0u);
                    my = tp2::ss_load64(ss, tp2::SS_MY); op = tp2::ss_load64(ss, tp2::SS_OP);
                    g.nn = 1u;
                    g.slow = p.no_reductions != 0u;
                    atomicAdd(&s_cnt[CNT_TREES], 1ull);
                    g.phase = PH_EXPLORE;
                } else {
                    go = false; // PH_DONE
                }
            }
            if (go && !err) {
                const uint32_t init = root.vis == 0.0f ? (uint32_t)tp2::K_INIT : 0u; // the construction visit (mcts.rs:133)
                if (FPUK == tp2::FPU_NORMAL_CACHED) err = tp2::descend<CW, tp2::FPU_NORMAL_CACHED, NT, PATH_CAP>(p, ss, g, root, my, op, pd, rc, path);
                else if (cfg.fpu_kind == SYN_FPU_CONST) err = tp2::descend<CW, SYN_FPU_CONST, NT, PATH_CAP>(p, ss, g, root, my, op, pd, rc, path);
                // <5, false> never sees Fpu Normal (engine.cu::launch_tpg_split sends it to the FPU_NORMAL_CACHED instantiation): one inlined descent less
                else if (cfg.fpu_kind == SYN_FPU_PARENT_Q || (FPUK == -1 && TEAMS == 5 && !PROF)) err = tp2::descend<CW, SYN_FPU_PARENT_Q, NT, PATH_CAP>(p, ss, g, root, my, op, pd, rc, path);
                else err = tp2::descend<CW, SYN_FPU_NORMAL, NT, PATH_CAP>(p, ss, g, root, my, op, pd, rc, path);
                pd.kind |= init;
            }
            if (err) { atomicCAS(p.error, 0, err); g.phase = PH_DONE; pd.kind = tp2::K_NONE; }
        }
        __syncwarp();
        { // statistics of the descent (summed per warp here so that they are not live across the forward)
            uint32_t a0 = __reduce_add_sync(0xffffffffu, rc.levels), a1 = __reduce_add_sync(0xffffffffu, rc.scanned);
            uint32_t a2 = __reduce_add_sync(0xffffffffu, rc.expansions), a3 = __reduce_add_sync(0xffffffffu, rc.created);
            if ((threadIdx.x & 31) == 0) { wc[tp2::WC_LEVELS] += a0; wc[tp2::WC_SCANNED] += a1; wc[tp2::WC_EXPANSIONS] += a2; wc[tp2::WC_CREATED] += a3; }
            rc.bp_levels = 0u; rc.leaf_evals = 0u;
        }
        long long t1 = PROF ? clock64() : 0;
        const bool need = (pd.kind & tp2::K_LEAF) != 0u;
        if (PROF) leaves += (uint32_t)__popc(__ballot_sync(0xffffffffu, need));
        if (!mlps::team_any(team, g.phase != PH_DONE)) break; // no thread of this team has a game left
        uint32_t mma_phase;
        const int slot = mlps::acquire_slot<TEAMS, SLOTS>(ms, team, r, mma_phase);
        long long t2 = PROF ? clock64() : 0;
        mlps::write_features<TEAMS, SLOTS>(ms, slot, r, my, op, need);
        float y[12];
        mlps::forward<TEAMS, SLOTS>(ms, p.mlp_bias, team, slot, r, mma_phase, y);
        mlps::release_slot<TEAMS, SLOTS>(ms, team, r, slot, mma_phase);
        long long t3 = PROF ? clock64() : 0;
        // ---- finish: child records for leaves, then ONE backprop site for every kind of explore
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
                tp2::write_children(g.nodes, pd, lg);
                v0 = __fdiv_rn(e0, tot); v1 = __fdiv_rn(e1, tot); v2 = __fdiv_rn(e2, tot);
                solved = (pd.lc >> 9) != 0u;
                rc.leaf_evals = 1u;
            } else {
                int idx = sol_index(pd.fc);
                v0 = idx == 0 ? 1.0f : 0.0f; v1 = idx == 1 ? 1.0f : 0.0f; v2 = idx == 2 ? 1.0f : 0.0f;
                solved = true;
            }
            rc.bp_levels = tp2::backprop<NT, PATH_CAP>(cfg, g.nodes, path, pd.depth, pd.id, v0, v1, v2, solved, g.slow);
            if (pd.kind & tp2::K_INIT) tp2::add_root_noise(p, ss, g.nodes);
        }
        __syncwarp();
        { // statistics of the finish: one shared-memory atomic per warp and counter
            uint32_t a4 = __reduce_add_sync(0xffffffffu, rc.bp_levels), a5 = __reduce_add_sync(0xffffffffu, rc.leaf_evals);
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

namespace eng {

// Batched Policy::eval (study-connect4/src/policies.rs:47-59) with the split forward: a CTA = one team of 128 positions.
__global__ void __launch_bounds__(128, 1) eval_split_kernel(const uint8_t* __restrict__ img_hi, const uint8_t* __restrict__ img_lo, const __grid_constant__ mlps::Bias bias,
                                                           const uint64_t* __restrict__ my_bb, const uint64_t* __restrict__ op_bb, uint32_t n,
                                                           float* __restrict__ logits, float* __restrict__ probs) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    mlps::Smem<1, 1>& ms = *reinterpret_cast<mlps::Smem<1, 1>*>(smem_raw);
    mlps::setup<1, 1>(ms, img_hi, img_lo);
    const int r = threadIdx.x;
    uint32_t mma_phase = ms.slot_phase[0];
    for (uint32_t base = blockIdx.x * 128u; base < n; base += gridDim.x * 128u) {
        const uint32_t idx = base + (uint32_t)r;
        const bool live = idx < n;
        const uint64_t my = live ? my_bb[idx] : 0ull, op = live ? op_bb[idx] : 0ull;
        mlps::write_features<1, 1>(ms, 0, r, my, op, live);
        float y[12];
        mlps::forward<1, 1>(ms, bias.b, 0, 0, r, mma_phase, y);
        if (live) {
            for (int k = 0; k < 9; ++k) logits[(size_t)idx * 9 + k] = y[k];
            float m = fmaxf(y[9], fmaxf(y[10], y[11]));
            float e0 = syn_expf(__fsub_rn(y[9], m)), e1 = syn_expf(__fsub_rn(y[10], m)), e2 = syn_expf(__fsub_rn(y[11], m));
            float tot = __fadd_rn(__fadd_rn(e0, e1), e2);
            probs[(size_t)idx * 3 + 0] = __fdiv_rn(e0, tot);
            probs[(size_t)idx * 3 + 1] = __fdiv_rn(e1, tot);
            probs[(size_t)idx * 3 + 2] = __fdiv_rn(e2, tot);
        }
        mlps::team_sync(0);
    }
    mlps::teardown<1, 1>(ms);
}

} // namespace eng
