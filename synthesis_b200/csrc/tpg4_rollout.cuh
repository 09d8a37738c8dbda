// tpg4_rollout.cuh — thread-per-game MCTS with RolloutPolicy leaves (BASELINE.json configs[0]) on tpg4_tree.cuh's
// family blocks: one 128-byte line per level of selection, memoised -q, the walk up's lines prefetched during the playout.
//
// Replaces synthesis/src/mcts.rs:29-489, synthesis/src/alpha_zero.rs:229-338 and
// synthesis/src/policies/rollout.rs:8-31 of the reference for rollout-only self-play.  The round structure, the
// ChaCha12 ring and the playout are tpg2_rollout.cuh's (see there); the tree code is tpg4_tree.cuh's.
#pragma once
#include "tpg2_rollout.cuh"
#include "tpg4.cuh"

namespace eng {

// One persistent CTA per SM, NT threads, thread = game (or search root); tpg4_tree.cuh's family blocks.
template <int NT, int FPU>
__global__ void __launch_bounds__(NT, 1) selfplay_rollout_tpg4_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ unsigned long long s_cnt[CNT_N];
    __shared__ unsigned long long s_wcnt[NT / 32][tp2::WC_N]; // per-warp statistics rows: plain adds by lane 0, no atomics
    constexpr int RING = tp2r::ring_words(NT), PATH_CAP = tp2r::path_cap(NT);
    uint32_t* const ring = reinterpret_cast<uint32_t*>(smem_raw) + threadIdx.x;      // [RING][NT]
    uint32_t* const path = reinterpret_cast<uint32_t*>(smem_raw) + RING * NT + threadIdx.x; // [PATH_CAP][NT]
    if (threadIdx.x < CNT_N) s_cnt[threadIdx.x] = 0ull;
    for (int i = threadIdx.x; i < (NT / 32) * tp2::WC_N; i += NT) (&s_wcnt[0][0])[i] = 0ull;
    unsigned long long* const wc = s_wcnt[threadIdx.x >> 5];
    __syncthreads();
    // a launch with fewer games than thread slots spreads them over all SMs, dealt round-robin to the CTA's warps
    const tp2::Seat seat = tp2::seat_of(p, 0, (int)threadIdx.x, NT / 32);
    const bool seated = seat.active;
    const size_t slot_id = seat.slot;
    uint32_t* const ss = p.slot_state + tp2::SS_WORDS * slot_id;
    const syn_mcts_cfg& cfg = p.cfg.mcts;
    const float stop_vis = (float)(p.cfg.num_explores + 1u); // explore_n is over when the root has 1 + num_explores visits
    struct { uint4* nodes; uint32_t nn, nb; int phase; } g;
    g.nodes = p.nodes + 2 * slot_id * p.arena_nodes;
    g.nn = 1u; g.nb = 1u; g.phase = seated ? PH_NEED_GAME : PH_DONE;
    const uint32_t arena_lines = p.arena_nodes >> 2;
    tp2r::RStream rs = {0u, 0u};
    for (;;) {
        // ---- cold bookkeeping, then at most one descent (as selfplay_nn_tpg2_kernel)
        tp2::Pend pd;
        pd.kind = tp2::K_NONE; pd.id = 0u; pd.fc = 0u; pd.lc = 0u; pd.depth = 0u;
        tp2::RoundCnt rc = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
        uint64_t my = 0, op = 0;
        if (g.phase != PH_DONE) {
            int err = 0;
            tp4::Cur root;
            bool go = true;
            if (g.phase == PH_EXPLORE) { // explore_n (mcts.rs:139-147): stop at num_explores or once the root is solved
                root = tp4::load_root(g.nodes);
                my = tp2::ss_load64(ss, tp2::SS_MY); op = tp2::ss_load64(ss, tp2::SS_OP);
                if (root.vis >= stop_vis || tp4::link_sol(root.link) != 0u) {
                    int pe = tp2::end_of_move(p, ss, g.nodes, g.nn, (uint32_t)root.vis - 1u, tp4::ReadRoot4{my | op});
                    g.phase = pe & 0xff; err = pe >> 8;
                    go = false; // the next tree starts next round
                }
            } else {
                if (g.phase == PH_NEED_GAME) {
                    g.phase = tp2::next_game(p, ss);
                    rs.rpos = 0u; rs.gen = 0u; // a new game (or search root) is a new stream
                }
                if (g.phase == PH_NEW_TREE) { // MCTS::with_capacity (mcts.rs:123-137): fresh arena, root only
                    root.id = tp4::ROOT_ID; root.vis = root.nq = 0.0f; root.link = 0u;
                    tp4::store_empty_root(g.nodes);
                    g.nb = 1u;
                    my = tp2::ss_load64(ss, tp2::SS_MY); op = tp2::ss_load64(ss, tp2::SS_OP);
                    g.nn = 1u;
                    atomicAdd(&s_cnt[CNT_TREES], 1ull);
                    g.phase = PH_EXPLORE;
                } else {
                    go = false; // PH_DONE
                }
            }
            if (go && !err) {
                const uint32_t init = root.vis == 0.0f ? (uint32_t)tp2::K_INIT : 0u; // the construction visit (mcts.rs:133)
                err = tp4::descend<FPU, NT, PATH_CAP>(cfg, arena_lines, tp2::FpuDraw{p, ss}, g.nodes, g.nb, g.nn, root, my, op, pd, rc, path);
                if (!err) tp4::prefetch_path<NT, PATH_CAP>(g.nodes, path, pd); // the walk up comes after the playout: its lines can travel meanwhile
                pd.kind |= init;
            }
            if (err) { atomicCAS(p.error, 0, err); g.phase = PH_DONE; pd.kind = tp2::K_NONE; }
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, g.phase != PH_DONE)) break; // no lane of this warp has a game left
        { // statistics of the descent
            uint32_t a0 = __reduce_add_sync(0xffffffffu, rc.levels), a1 = __reduce_add_sync(0xffffffffu, rc.scanned);
            uint32_t a2 = __reduce_add_sync(0xffffffffu, rc.expansions), a3 = __reduce_add_sync(0xffffffffu, rc.created);
            if ((threadIdx.x & 31) == 0) { wc[tp2::WC_LEVELS] += a0; wc[tp2::WC_SCANNED] += a1; wc[tp2::WC_EXPANSIONS] += a2; wc[tp2::WC_CREATED] += a3; }
        }
        const bool need = (pd.kind & tp2::K_LEAF) != 0u;
        // ---- RolloutPolicy::eval: top up the stream at one place, then play the position out
        uint32_t plies = 0u;
        int idx = 0;
        uint64_t seed = 0ull;
        if (need) {
            seed = tp2::stream_seed(p, ss[tp2::SS_GI], SYN_STREAM_ROLLOUT);
            while (rs.gen - rs.rpos <= (uint32_t)(RING - 16)) {
                tp2r::gen_block<NT>(seed, rs.gen >> 4, ring, rs.gen & (uint32_t)(RING - 1));
                rs.gen += 16u;
            }
        }
        __syncwarp();
        if (need) idx = tp2r::rollout<NT, RING>(seed, rs, ring, my, op, plies);
        __syncwarp();
        // ---- finish: child records for leaves, then ONE backprop site for every kind of explore
        uint32_t bp_levels = 0u;
        if (pd.kind != tp2::K_NONE) {
            float v0, v1, v2;
            bool solved;
            if (need) {
                tp4::write_children_uniform(g.nodes, pd);
                v0 = idx == 0 ? 1.0f : 0.0f; v1 = idx == 1 ? 1.0f : 0.0f; v2 = idx == 2 ? 1.0f : 0.0f;
                solved = (pd.lc >> 9) != 0u;
            } else {
                int si = sol_index(pd.fc);
                v0 = si == 0 ? 1.0f : 0.0f; v1 = si == 1 ? 1.0f : 0.0f; v2 = si == 2 ? 1.0f : 0.0f;
                solved = true;
            }
            tp4::backprop<NT, PATH_CAP, tp4::BP_CHUNK>(cfg, g.nodes, path, pd.depth, pd.id, v0, v1, v2, solved);
            bp_levels = pd.depth + 1u; // mcts.rs:431: every level from the leaf to the root
            if (pd.kind & tp2::K_INIT) tp4::add_root_noise(p, ss, g.nodes);
        }
        __syncwarp();
        { // statistics of the finish: one shared-memory atomic per warp and counter
            uint32_t a4 = __reduce_add_sync(0xffffffffu, bp_levels), a5 = __reduce_add_sync(0xffffffffu, need ? 1u : 0u);
            uint32_t a6 = __reduce_add_sync(0xffffffffu, plies);
            if ((threadIdx.x & 31) == 0) { wc[tp2::WC_BACKPROP] += a4; wc[tp2::WC_LEAF_EVALS] += a5; wc[tp2::WC_ROLLOUT_PLIES] += a6; }
        }
    }
    __syncthreads();
    if (threadIdx.x < CNT_N && s_cnt[threadIdx.x]) atomicAdd(p.counters + threadIdx.x, s_cnt[threadIdx.x]);
    if (threadIdx.x < tp2::WC_N) {
        unsigned long long t = 0ull;
        for (int w = 0; w < NT / 32; ++w) t += s_wcnt[w][threadIdx.x];
        if (t) atomicAdd(p.counters + tp2::wc_counter(threadIdx.x), t);
    }
}

} // namespace eng
