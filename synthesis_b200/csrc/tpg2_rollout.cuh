// tpg2_rollout.cuh — thread-per-game MCTS with RolloutPolicy leaves (BASELINE.json configs[0]).
//
// Replaces synthesis/src/mcts.rs:29-489, synthesis/src/alpha_zero.rs:229-338 and
// synthesis/src/policies/rollout.rs:8-31 of the reference for rollout-only self-play.  Same tree code as
// tpg2.cuh (descend / backprop / end_of_move: one explore per thread per round, node records of 32 bytes,
// backprop by vector reductions at the L2), but the leaf is evaluated by the thread itself: a random
// playout on the game's own ChaCha12 stream (StdRng, one u32 per ply + rejections, rollout.rs:16).
//
// What shapes the schedule (the lane-group kernel selfplay_rollout_kernel<16> idles 15 of 16 lanes while
// one lane plays the rollout, 151 M explores/s):
//  * a warp's lanes are 32 different games; every phase of a round (descend, top up the stream, rollout,
//    finish) is entered by all lanes together, so divergence stays inside a phase.
//  * ChaCha12 blocks are generated at ONE place per round, before the rollout, into a per-thread ring of
//    RING words in shared memory (word-major, stride NT: conflict-free whatever each lane's position is).
//    The ring is topped up to more than RING-16 words, so a rollout almost never runs dry; when it does the
//    block is generated in the loop (same function, cold).  Without the ring every lane would refill at a
//    different ply and the 600-instruction block function would run with one or two lanes active in most
//    iterations.
//  * the key is not kept: seed_from_u64's PCG32 expansion (8 steps) is redone per block from the stream
//    seed, which costs 1/8 of the block itself and frees 8 registers / 32 bytes of shared memory per game.
//  * no shared-memory weights and no team barrier: NT threads per CTA is only limited by registers.
#pragma once
#include "tpg2.cuh"

namespace tp2r {

using namespace eng;

// One ChaCha12 block of the stream `seed`, block index `blk`, written word-major into ring[(w0 + i) * NT].
template <int NT>
__device__ __noinline__ void gen_block(uint64_t seed, uint32_t blk, uint32_t* ring, uint32_t w0) {
    uint32_t key[8];
    rng::seed_key(seed, key);
    uint32_t x0 = 0x61707865u, x1 = 0x3320646eu, x2 = 0x79622d32u, x3 = 0x6b206574u;
    uint32_t x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3], x8 = key[4], x9 = key[5], x10 = key[6], x11 = key[7];
    uint32_t x12 = blk, x13 = 0u, x14 = 0u, x15 = 0u;
#pragma unroll 1
    for (int r = 0; r < 6; ++r) {
        SYN_QR(x0, x4, x8, x12) SYN_QR(x1, x5, x9, x13) SYN_QR(x2, x6, x10, x14) SYN_QR(x3, x7, x11, x15)
        SYN_QR(x0, x5, x10, x15) SYN_QR(x1, x6, x11, x12) SYN_QR(x2, x7, x8, x13) SYN_QR(x3, x4, x9, x14)
    }
    uint32_t* o = ring + (size_t)w0 * NT;
    o[0 * NT] = x0 + 0x61707865u; o[1 * NT] = x1 + 0x3320646eu; o[2 * NT] = x2 + 0x79622d32u; o[3 * NT] = x3 + 0x6b206574u;
    o[4 * NT] = x4 + key[0]; o[5 * NT] = x5 + key[1]; o[6 * NT] = x6 + key[2]; o[7 * NT] = x7 + key[3];
    o[8 * NT] = x8 + key[4]; o[9 * NT] = x9 + key[5]; o[10 * NT] = x10 + key[6]; o[11 * NT] = x11 + key[7];
    o[12 * NT] = x12 + blk; o[13 * NT] = x13; o[14 * NT] = x14; o[15 * NT] = x15;
}

// The game's rollout stream as the thread sees it: words [rpos, gen) are in the ring.
struct RStream {
    uint32_t rpos, gen; // stream words consumed / generated (gen is a multiple of 16)
};

// rand 0.8 gen_range(0..n) zone for n in 1..9: 0xffffffff - (2^32 mod n), 2^32 mod n packed 4 bits per n.
__device__ __forceinline__ uint32_t gen_range_zone(uint32_t n) { return 0xffffffffu - (uint32_t)((0x4044101000ull >> (4u * n)) & 15ull); }

// policies/rollout.rs:8-31 from a non-terminal leaf.  Returns the one-hot index (0 Lose, 1 Draw, 2 Win) for
// the leaf's player to move.
template <int NT, int RING>
__device__ __forceinline__ int rollout(uint64_t seed, RStream& rs, uint32_t* ring, uint64_t my, uint64_t op, uint32_t& plies) {
    uint32_t k = 0;
    for (;;) {
        const uint64_t occ = my | op;
        const uint64_t legal = (~(occ >> 6)) & c4::ROW0; // bit 7c set <=> column c has room
        const uint32_t llo = (uint32_t)legal, lhi = (uint32_t)(legal >> 32);
        const uint32_t nlo = (uint32_t)__popc(llo), n = nlo + (uint32_t)__popc(lhi);
        const uint32_t zone = gen_range_zone(n);
        uint32_t hi;
        for (;;) {
            if (rs.gen == rs.rpos) { // ran dry inside a rollout: cold
                gen_block<NT>(seed, rs.gen >> 4, ring, rs.gen & (uint32_t)(RING - 1));
                rs.gen += 16u;
            }
            uint32_t v = ring[(size_t)(rs.rpos & (uint32_t)(RING - 1)) * NT];
            ++rs.rpos;
            uint64_t m = (uint64_t)v * (uint64_t)n;
            hi = (uint32_t)(m >> 32);
            if ((uint32_t)m <= zone) break;
        }
        // hi-th legal column, ascending: at most 4 bits to clear in one 32-bit half
        const bool low = hi < nlo;
        uint32_t mask = low ? llo : lhi;
        const uint32_t t = low ? hi : hi - nlo;
#pragma unroll
        for (uint32_t i = 0; i < 4u; ++i)
            if (i < t) mask &= mask - 1u;
        const int p7 = (low ? 0 : 32) + __ffs((int)mask) - 1;
        const uint64_t bit = (occ + (1ull << p7)) & (0x7full << p7);
        const uint64_t mover = my | bit;
        my = op;
        op = mover;
        ++k;
        if (c4::won(mover)) { plies += k; return (k & 1u) ? 2 : 0; }
        if ((occ | bit) == c4::ALL) { plies += k; return 1; }
    }
}

// visit() after RolloutPolicy::eval: all logits are 0, so the stable softmax is exp(0) = 1 per legal child over a
// total of n (mcts.rs:409-423) — the same bits as tree.cuh's explore_finish(uniform_prior).
__device__ __forceinline__ void write_children_uniform(uint4* nodes, const tp2::Pend& pd) {
    const uint32_t legal = pd.lc & 0x1ffu, csol2 = pd.lc >> 9;
    const uint32_t prior = __float_as_uint(__fdiv_rn(1.0f, (float)__popc(legal)));
    uint32_t rank = 0u;
#pragma unroll
    for (int col = 0; col < 9; ++col) {
        if ((legal >> col) & 1u) {
            uint32_t s2 = (csol2 >> (2 * col)) & 3u;
            uint32_t csol = s2 == 1u ? c4::SOL_LOSE0 : (s2 == 2u ? c4::SOL_DRAW0 : 0u);
            tp2::store_rec(nodes, pd.fc + rank, 0.f, 0.f, 0.f, 0.f, prior, pd.id, 0u, (csol << 8) | ((uint32_t)col << 16));
            ++rank;
        }
    }
}

// 32 words (two blocks) per game at every thread count: a 64-word ring tops up less often, but the shared memory it takes
// comes out of the L1 (the SM's carve-out goes up in steps), where this kernel keeps its spills and hottest nodes
// (profiles/r1_rollout_keys_smem_ab.txt, r1_rollout_variants.txt).
__host__ __device__ constexpr int ring_words(int nt) { return 32; }
__host__ __device__ constexpr int path_cap(int nt) { return nt <= 512 ? 12 : 8; }
__host__ __device__ constexpr size_t smem_bytes(int nt) { return (size_t)nt * (size_t)(ring_words(nt) + path_cap(nt)) * sizeof(uint32_t); }

} // namespace tp2r

namespace eng {

// One persistent CTA per SM, NT threads, thread = game (or search root).  CW = child records per memory round trip.
template <int NT, int CW, int FPUK = -1> // FPUK: see selfplay_nn_tpg2_kernel
__global__ void __launch_bounds__(NT, 1) selfplay_rollout_tpg2_kernel(const __grid_constant__ KParams p) {
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
    const tp2::Seat seat = tp2::seat_of(p, 0, (int)threadIdx.x, NT / 32); // few games: spread over all SMs and warps
    const size_t slot_id = seat.slot;
    uint32_t* const ss = p.slot_state + tp2::SS_WORDS * slot_id;
    const syn_mcts_cfg& cfg = p.cfg.mcts;
    const float stop_vis = (float)(p.cfg.num_explores + 1u); // explore_n is over when the root has 1 + num_explores visits
    tp2::Game g;
    g.nodes = p.nodes + 2 * slot_id * p.arena_nodes;
    g.nn = 1u; g.phase = seat.active ? PH_NEED_GAME : PH_DONE; g.slow = p.no_reductions != 0u;
    tp2r::RStream rs = {0u, 0u};
    for (;;) {
        // ---- cold bookkeeping, then at most one descent (as selfplay_nn_tpg2_kernel)
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
                if (g.phase == PH_NEED_GAME) {
                    g.phase = tp2::next_game(p, ss);
                    rs.rpos = 0u; rs.gen = 0u; // a new game (or search root) is a new stream
                }
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
                else if (FPUK == SYN_FPU_CONST) err = tp2::descend<CW, SYN_FPU_CONST, NT, PATH_CAP>(p, ss, g, root, my, op, pd, rc, path); // one descent only: 80 bytes of spills instead of 104, +3.5 %
                else if (cfg.fpu_kind == SYN_FPU_CONST) err = tp2::descend<CW, SYN_FPU_CONST, NT, PATH_CAP>(p, ss, g, root, my, op, pd, rc, path);
                // <1024, 3> never sees Fpu Normal (engine.cu::launch_rollout_tpg sends it to the FPU_NORMAL_CACHED instantiation): one inlined descent less
                else if (cfg.fpu_kind == SYN_FPU_PARENT_Q || (FPUK == -1 && NT == 1024 && CW == 3)) err = tp2::descend<CW, SYN_FPU_PARENT_Q, NT, PATH_CAP>(p, ss, g, root, my, op, pd, rc, path);
                else err = tp2::descend<CW, SYN_FPU_NORMAL, NT, PATH_CAP>(p, ss, g, root, my, op, pd, rc, path);
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
                tp2r::write_children_uniform(g.nodes, pd);
                v0 = idx == 0 ? 1.0f : 0.0f; v1 = idx == 1 ? 1.0f : 0.0f; v2 = idx == 2 ? 1.0f : 0.0f;
                solved = (pd.lc >> 9) != 0u;
            } else {
                int si = sol_index(pd.fc);
                v0 = si == 0 ? 1.0f : 0.0f; v1 = si == 1 ? 1.0f : 0.0f; v2 = si == 2 ? 1.0f : 0.0f;
                solved = true;
            }
            bp_levels = tp2::backprop<NT, PATH_CAP>(cfg, g.nodes, path, pd.depth, pd.id, v0, v1, v2, solved, g.slow);
            if (pd.kind & tp2::K_INIT) tp2::add_root_noise(p, ss, g.nodes);
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
