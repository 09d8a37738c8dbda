// tpg2.cuh — thread-per-game MCTS, round-synchronous form: ONE explore per thread per round.
//
// Replaces synthesis/src/mcts.rs:29-489 and synthesis/src/alpha_zero.rs:229-338 of the
// reference, one THREAD per game, like tpg.cuh (same algorithm, same f32 operation order, same
// results bit for bit).  What differs is the schedule, shaped by what ncu showed on tpg.cuh
// (profiles/r1d_*): 12 of 32 lanes active per instruction, 6.3 % of all instructions register
// spills whose local-memory sectors out-numbered the node loads, and a quarter of all stall samples
// at the team barrier waiting for threads that ran a second descent.
//
//  * a round is exactly: [cold tree/game bookkeeping] -> descend -> (team forward) -> finish.
//    An explore that ends on a proven node does NOT loop into another descent: it simply has no
//    leaf in this round's tile, so all lanes of a warp walk the same code once per round.
//  * backprop happens at ONE place for every kind of explore (leaf, proven node, auto-extended
//    terminal), after the forward, so the lanes of a warp run it together.
//  * per-thread state that is only touched once per move (game index, ply, stream positions)
//    lives in a 32-byte slot record in global memory, and the statistics counters are summed per
//    warp (redux) into shared memory — the hot loop keeps ~12 registers of state instead of ~35.
//  * the root record is loaded once per round and serves both the explore_n stop test
//    (mcts.rs:139-147) and the first level of selection.
//
// Node record (32 bytes = one sector), words: 0 num_visits | 1..3 outcome sums L,D,W |
// 4 action_prob | 5 parent | 6 first_child | 7 num_children | solution<<8 | action<<16.
#pragma once
#ifndef SYN_TPG2_DISPATCH
#define SYN_TPG2_DISPATCH 1 // 1 = no Normal descent in <5,.,false> (a Const-only instantiation was measured too: +0.6 % instead of +1.4 %); 0 = three-way
#endif
#ifndef SYN_TPG2_CBIAS
#define SYN_TPG2_CBIAS 0
#endif
#include "mlp_team.cuh"
#include "selfplay.cuh"

namespace tp2 {

using namespace eng;
using tpx::K_NONE; using tpx::K_LEAF; using tpx::K_TERMINAL; using tpx::K_INIT;
using tpx::Pend; using tpx::RoundCnt; using tpx::RootOut;

struct Rec {
    float vis, o0, o1, o2;
    uint32_t prior, parent, fc, pk;
};

__device__ __forceinline__ Rec load_rec(const uint4* nodes, uint32_t i) { // LDG.E.256: one sector, one request
    Rec r;
    unsigned long long q0, q1, q2, q3;
    asm volatile("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(nodes + 2 * (size_t)i) : "memory");
    r.vis = __uint_as_float((uint32_t)q0); r.o0 = __uint_as_float((uint32_t)(q0 >> 32));
    r.o1 = __uint_as_float((uint32_t)q1); r.o2 = __uint_as_float((uint32_t)(q1 >> 32));
    r.prior = (uint32_t)q2; r.parent = (uint32_t)(q2 >> 32); r.fc = (uint32_t)q3; r.pk = (uint32_t)(q3 >> 32);
    return r;
}
__device__ __forceinline__ void store_rec(uint4* nodes, uint32_t i, float vis, float o0, float o1, float o2, uint32_t prior,
                                          uint32_t parent, uint32_t fc, uint32_t pk) { // STG.E.256
    unsigned long long q0 = (unsigned long long)__float_as_uint(vis) | ((unsigned long long)__float_as_uint(o0) << 32);
    unsigned long long q1 = (unsigned long long)__float_as_uint(o1) | ((unsigned long long)__float_as_uint(o2) << 32);
    unsigned long long q2 = (unsigned long long)prior | ((unsigned long long)parent << 32);
    unsigned long long q3 = (unsigned long long)fc | ((unsigned long long)pk << 32);
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(nodes + 2 * (size_t)i), "l"(q0), "l"(q1), "l"(q2), "l"(q3) : "memory");
}
__device__ __forceinline__ void store_stat(uint4* nodes, uint32_t i, float vis, float o0, float o1, float o2) {
    nodes[2 * (size_t)i] = make_uint4(__float_as_uint(vis), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2));
}
// Fire-and-forget form of store_stat(load + add): ONE vector reduction at the L2 (REDG.E.ADD.F32x4.RN), no
// round trip.  Each component is an IEEE round-to-nearest-even f32 addition like the CPU's, EXCEPT that the
// L2's adder flushes subnormal inputs and results to zero — red_exact() below says when that cannot happen.
__device__ __forceinline__ void red_stat(uint4* nodes, uint32_t i, float v0, float v1, float v2) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(nodes + 2 * (size_t)i), "f"(1.0f), "f"(v0), "f"(v1), "f"(v2) : "memory");
}
// True when x is zero or |x| >= 2^-100.  As long as every value ever added into a tree passes this test,
// every outcome sum in the tree is zero or normal (sums and differences of such values are zero or >= 2^-124),
// so the flushing adder and the IEEE adder agree bit for bit.  The first value that fails makes the tree
// "slow" (Game::slow) until it is reset: it then only uses load / FADD / store.
__device__ __forceinline__ bool red_exact(float x) { return ((__float_as_uint(x) << 1) >= (27u << 24)) || ((__float_as_uint(x) << 1) == 0u); }
__device__ __forceinline__ void prefetch_l2(const uint4* nodes, uint32_t i) { asm volatile("prefetch.global.L2 [%0];" ::"l"(nodes + 2 * (size_t)i)); }
__device__ __forceinline__ uint32_t* meta_words(uint4* nodes, uint32_t i) { return reinterpret_cast<uint32_t*>(nodes + 2 * (size_t)i + 1); }
enum { MW_PRIOR = 0, MW_PARENT = 1, MW_FC = 2, MW_PK = 3 };

// Per-slot state outside the registers (global memory, 64 bytes per game in flight): the root position of
// the current tree (read once per round next to the root record) and what only changes once per move.
enum { SS_MY = 0, SS_OP = 2, SS_GI = 4, SS_PLY = 5, SS_APOS = 6, SS_FPU_POS = 7, SS_NOISE_POS = 8, SS_WORDS = 16 };
// A second record per slot, touched only by the shipped Normal first-play urgency (fpu_normal_draw): the FPU stream's
// expanded key, its current ChaCha12 block and which block that is.  (Kept out of the slot record above: growing that one
// from 64 to 192 bytes cost the Fpu::Const bench 4 %.  The stream position stays in the slot record: position 0 = the
// game's first draw = the moment to expand its key.)
enum { FS_KEY = 0, FS_BLOCK = 8, FS_BLK = 24, FS_WORDS = 32 };


// Hot per-thread state is three registers: the arena pointer is recomputed from the slot index, and
// the explore count of the current tree IS the root's visit count (every backprop ends at the root:
// num_visits(root) == 1 + explores done, mcts.rs:133, 480), so neither needs a register.
struct Game {
    uint4* nodes;
    uint32_t nn; // nodes.len()
    int phase;   // PH_*
    bool slow;   // this tree has seen a value the L2's flushing adder would treat differently: no reductions (red_exact)
};


// Levels of the path table per thread: what fits beside the MLP state in 227 KB of shared memory.
__host__ __device__ constexpr int path_cap(int teams) { return teams <= 4 ? 12 : teams == 5 ? 10 : teams == 6 ? 8 : 6; }

// Per-warp statistics rows in shared memory (see the kernels) and the global counter each column belongs to.
enum { WC_LEVELS = 0, WC_SCANNED, WC_EXPANSIONS, WC_CREATED, WC_BACKPROP, WC_LEAF_EVALS, WC_ROLLOUT_PLIES, WC_N };
__device__ __forceinline__ int wc_counter(int i) {
    return i == WC_LEVELS ? CNT_SELECT_LEVELS : i == WC_SCANNED ? CNT_CHILDREN_SCANNED : i == WC_EXPANSIONS ? CNT_EXPANSIONS
           : i == WC_CREATED ? CNT_CHILDREN_CREATED : i == WC_BACKPROP ? CNT_BACKPROP_LEVELS : i == WC_LEAF_EVALS ? CNT_LEAF_EVALS : CNT_ROLLOUT_PLIES;
}


__device__ __forceinline__ uint64_t stream_seed(const KParams& p, uint32_t gi, unsigned k) {
    if (!p.search_mode) return syn_stream_seed(p.seed, p.first_game + gi, k);
    uint64_t s = p.pos_seed[gi];
    if (k == SYN_STREAM_ROLLOUT) return s;
    if (k == SYN_STREAM_ACTION) return 0ull;
    return (s ^ (1ull << 63)) + (k == SYN_STREAM_FPU ? 1ull : 0ull);
}

// mcts.rs:354 with the shipped closure (study-connect4/src/main.rs:43-47).  The closure runs for EVERY unvisited child EVERY
// time its parent is selected through, so with the shipped configuration this is on the hot path: re-deriving the stream's
// key (PCG32 x 8) and its ChaCha12 block (~700 instructions) per draw made the thread-per-game kernels 4.6 x slower than with
// a constant FPU (265 M against 1,232 M explores/s, profiles/r2_normal_fpu.txt).  Key and current block now live in a
// 128-byte record per game in flight (KParams::fpu_state, L2-resident) and a block is generated once per 16 words.
__device__ __noinline__ float fpu_normal_draw(const KParams& p, uint32_t* ss) { // mcts.rs:354 with the shipped closure; cold
    rng::Stream st;
    st.init(stream_seed(p, ss[SS_GI], SYN_STREAM_FPU), ss[SS_FPU_POS]);
    float v = syn_normal(st, p.cfg.mcts.fpu_a, p.cfg.mcts.fpu_b);
    ss[SS_FPU_POS] = (uint32_t)st.pos;
    return v;
}

// The same draw with the stream's expanded key and current ChaCha12 block cached per game in flight (KParams::fpu_state, 128 bytes
// per slot, L2-resident; position 0 = the game's first draw = the moment to expand its key): a block is generated once per 16
// words instead of once per draw — 265 -> ~500 M explores/s with NN leaves at full scale (profiles/r2_normal_fpu.txt).
// Only the Normal-only kernel instantiations (FPUK = FPU_NORMAL_CACHED) call it, and it has its own block function: the tuned
// runtime-dispatch kernels must not see a changed callee — giving rng::chacha12_block one new caller moved the register
// allocation of selfplay_nn_tpg2_kernel<5,4> and cost the Fpu::Const bench 2.5 % (found by diffing SASS between builds).
constexpr int FPU_NORMAL_CACHED = 100 + SYN_FPU_NORMAL;
struct SlotFpuStream { // R-concept of include/syn_sampling.h
    uint32_t* fs;
    uint32_t pos, have_blk;
    __device__ uint32_t next_u32() {
        const uint32_t blk = pos >> 4;
        if (blk != have_blk) {
            chacha12_block_smem(fs, blk); // key at fs[0, 8), block to fs[8, 24): the layout of the lane-group kernels' copy (tree.cuh)
            fs[FS_BLK] = blk;
            have_blk = blk;
        }
        return fs[FS_BLOCK + ((pos++) & 15u)];
    }
};
__device__ __noinline__ float fpu_normal_draw_cached(const KParams& p, uint32_t* ss) {
    uint32_t* fs = p.fpu_state + (size_t)FS_WORDS * (size_t)((ss - p.slot_state) / SS_WORDS);
    const uint32_t pos = ss[SS_FPU_POS];
    uint32_t have = fs[FS_BLK];
    if (pos == 0u) { // first draw of this game (next_game reset the position): expand the key once
        uint32_t key[8];
        rng::seed_key(stream_seed(p, ss[SS_GI], SYN_STREAM_FPU), key);
#pragma unroll
        for (int i = 0; i < 8; ++i) fs[FS_KEY + i] = key[i];
        have = 0xffffffffu;
    }
    SlotFpuStream st{fs, pos, have};
    const float v = syn_normal(st, p.cfg.mcts.fpu_a, p.cfg.mcts.fpu_b);
    ss[SS_FPU_POS] = st.pos;
    return v;
}

// Which arena slot a thread of a team-structured CTA plays in, and whether it plays at all: a launch that holds fewer
// games than the GPU has thread slots spreads them over ALL SMs instead of filling the first CTAs (the reference plays
// 1,000 games per iteration, study-connect4/src/main.rs:26: 6-7 per SM instead of 640 on two).  CTA b seats
// seats_q + (b < seats_rem) games; seat order runs round-robin over the teams in use and, inside a team, over its
// warps (seat li of a team = lane li / W of warp li % W), so that the warps of a CTA carry equal numbers of games.
struct Seat { bool active; size_t slot; };
__device__ __forceinline__ Seat seat_of(const KParams& p, int team, int r, int warps_per_team) {
    const uint32_t li = (uint32_t)(r & 31) * (uint32_t)warps_per_team + (uint32_t)(r >> 5);
    const uint32_t n_b = p.seats_q + (blockIdx.x < p.seats_rem ? 1u : 0u);
    Seat s;
    s.active = (uint32_t)team < p.teams_used && li < p.per_team && li * p.teams_used + (uint32_t)team < n_b;
    s.slot = ((size_t)blockIdx.x * p.teams_used + (size_t)team) * p.per_team + li;
    return s;
}

// mcts.rs:354 with the shipped closure, on the game's seeded FPU stream (a functor for tpg4_tree.cuh's descend)
struct FpuDraw {
    const KParams& p;
    uint32_t* ss;
    __device__ __forceinline__ float operator()() const { return fpu_normal_draw(p, ss); }
};

// One explore from the (already loaded) root up to the point where the policy is needed
// (mcts.rs:310-325, 327-372, 374-406).  `my`/`op` enter as the root position and leave as the
// leaf's.  Returns an error code (0 = none).
// CW = child records requested per memory round trip (3 at 64 registers per thread, 9 at 128).
// FPU = the configured syn_fpu_kind, a template parameter so that the common Fpu::Const instantiation carries
// neither the parent's outcome sums (ParentQ) nor a call in its inner loop (Normal).
// path = this thread's column of the CTA's path table in shared memory (entry l-1 = the node walked at level l,
// stride NT words, levels 1 .. PATH_CAP), so that backprop knows the way up without reading parent links.
template <int CW, int FPU, int NT, int PATH_CAP>
__device__ __forceinline__ int descend(const KParams& p, uint32_t* ss, Game& g, const Rec& root, uint64_t& my, uint64_t& op, Pend& pd,
                                       RoundCnt& rc, uint32_t* path) {
    const syn_mcts_cfg& cfg = p.cfg.mcts;
    uint4* nodes = g.nodes;
    uint32_t cur = 0u;
    constexpr bool PQ = FPU == SYN_FPU_PARENT_Q;
    float cvis = root.vis, cop0 = PQ ? root.o0 : 0.0f, cop2 = PQ ? root.o2 : 0.0f;
    uint32_t cfc = root.fc, cpk = root.pk;
    uint32_t depth = 0;
    const bool puct = cfg.exploration_kind == SYN_EXPLORATION_POLYNOMIAL_UCT;
    for (;;) {
        uint32_t sol = (cpk >> 8) & 0xffu, nch = cpk & 0xffu;
        if (sol) { rc.levels = depth; pd.kind = K_TERMINAL; pd.id = cur; pd.fc = sol; pd.depth = depth; return 0; } // mcts.rs:314-316
        if (nch == 0u) break;
        // the children are read CW at a time; DRAM hands out whole 128-byte lines, so asking for the last child's line now
        // makes the second batch an L2 hit instead of a second trip to HBM
        if (nch > (uint32_t)CW) prefetch_l2(nodes, cfc + nch - 1u);
        // three batches (CW = 3, seven or more children): nine records span three or four lines, so the middle one is asked for too
        if (CW < 5 && nch > 2u * (uint32_t)CW) prefetch_l2(nodes, cfc + (uint32_t)CW + 1u);
        // ---- select_best_child (mcts.rs:327-372): first strict maximum in child order
        const float pterm = puct ? __fsqrt_rn(cvis) : __fsqrt_rn(__fmul_rn(cfg.c, syn_logf(cvis)));
        const float fpu_q = PQ ? __fdiv_rn(__fsub_rn(cop2, cop0), cvis) : cfg.fpu_a; // Fpu::ParentQ = parent.q() (mcts.rs:353), once per level
        uint32_t b = 0u, bfc = 0u, bpk = 0u;
        float bval = 0.0f, bvis = 0.0f, bo0 = 0.0f, bo2 = 0.0f;
        for (uint32_t k0 = 0; k0 < nch; k0 += (uint32_t)CW) { // CW records per memory round trip
            Rec chs[CW];
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)CW; ++j) chs[j] = load_rec(nodes, cfc + min(k0 + j, nch - 1u));
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)CW; ++j) {
                const uint32_t k = k0 + j;
                const Rec& ch = chs[j];
                uint32_t csol = (ch.pk >> 8) & 0xffu, cn = ch.pk & 0xffu;
                float q;
                // (a branch-free form — all three candidates computed and selected — measured 3 % slower: the extra
                // divisions cost more than the divergence they remove)
                if (csol) {
                    uint32_t kd = sol_kind(csol);
                    q = cfg.select_solved_nodes ? (kd == SYN_KIND_WIN ? -1.0f : (kd == SYN_KIND_LOSE ? 1.0f : 0.0f)) : __uint_as_float(0xff800000u);
                } else if (cn == 0u) {
                    if (FPU == SYN_FPU_NORMAL) q = k < nch ? fpu_normal_draw(p, ss) : 0.0f;
                    else if (FPU == FPU_NORMAL_CACHED) q = k < nch ? fpu_normal_draw_cached(p, ss) : 0.0f;
                    else q = fpu_q;
                } else {
                    q = -__fdiv_rn(__fsub_rn(ch.o2, ch.o0), ch.vis);
                }
                float u = puct ? __fdiv_rn(__fmul_rn(__fmul_rn(cfg.c, __uint_as_float(ch.prior)), pterm), __fadd_rn(1.0f, ch.vis))
                               : __fdiv_rn(pterm, __fsqrt_rn(ch.vis));
                float value = __fadd_rn(q, u);
                if (k < nch && (k == 0u || value > bval)) {
                    b = k; bval = value; bvis = ch.vis; bfc = ch.fc; bpk = ch.pk;
                    if (PQ) { bo0 = ch.o0; bo2 = ch.o2; }
                }
            }
        }
        rc.scanned += nch;
        cur = cfc + b;
        cvis = bvis; cop0 = bo0; cop2 = bo2; cfc = bfc; cpk = bpk;
        c4::step(my, op, (int)((cpk >> 16) & 0xffu));
        if (depth < (uint32_t)PATH_CAP) path[depth * NT] = cur;
        if (++depth >= 64u) return DERR_DEPTH_OVERFLOW;
    }
    rc.levels = depth; // one select_best_child call per level walked
    // ---- visit (mcts.rs:374-406): number the children of `cur`; auto-extend through only-children
    for (;;) {
        uint64_t occ = my | op;
        const uint64_t win = c4::winning_cells(my);   // where the mover completes four in a row
        const bool last = __popcll(occ) == 62;        // this move fills the board: a draw unless it wins
        uint32_t lm = 0u, cs2 = 0u, n = 0u;
#pragma unroll
        for (int col = 0; col < 9; ++col) {
            uint32_t colbits = (uint32_t)((occ >> (7 * col)) & 0x7full);
            if (colbits != 0x7fu) {
                uint64_t bit = 1ull << (7 * col + __popc(colbits));
                uint32_t s2 = (win & bit) ? 1u : (last ? 2u : 0u);
                lm |= 1u << col;
                cs2 |= s2 << (2 * col);
                ++n;
            }
        }
        uint32_t fc = g.nn;
        if (fc + n > p.arena_nodes) return DERR_ARENA_OVERFLOW;
        uint32_t* mw = meta_words(nodes, cur); // mark_visited (mcts.rs:399-400): fc and pk are adjacent words
        *reinterpret_cast<uint2*>(mw + MW_FC) = make_uint2(fc, (cpk & 0xffffff00u) | n);
        g.nn = fc + n;
        rc.expansions += 1u;
        rc.created += n;
        if (cfg.auto_extend && n == 1u) { // mcts.rs:404-405: recurse into the only child, no policy call
            int only = __ffs(lm) - 1;
            uint32_t s2 = (cs2 >> (2 * only)) & 3u;
            uint32_t osol = s2 == 1u ? c4::SOL_LOSE0 : (s2 == 2u ? c4::SOL_DRAW0 : 0u);
            cpk = (osol << 8) | ((uint32_t)only << 16);
            store_rec(nodes, fc, 0.f, 0.f, 0.f, 0.f, __float_as_uint(1.0f), cur, 0u, cpk);
            c4::step(my, op, only);
            cur = fc;
            if (depth < (uint32_t)PATH_CAP) path[depth * NT] = cur;
            if (++depth >= 64u) return DERR_DEPTH_OVERFLOW;
            if (osol) { pd.kind = K_TERMINAL; pd.id = cur; pd.fc = osol; pd.depth = depth; return 0; } // mcts.rs:377-379
            continue;
        }
        pd.kind = K_LEAF; pd.id = cur; pd.fc = fc; pd.lc = lm | (cs2 << 9); pd.depth = depth;
        return 0;
    }
}

// mcts.rs:429-488 from node `id` (at level `depth`) up to the root.  While the value is still "solved" the node and its
// children are read like the reference does; once it is not (the common case from the first level on), a level is ONE
// vector reduction into the node's {visits, outcome sums} and the way up comes from the path table, so nothing waits for
// memory.  Beyond PATH_CAP levels the parent link is read.  `slow` (Game::slow) forces load / add / store everywhere.
template <int NT, int PATH_CAP>
__device__ __forceinline__ uint32_t backprop(const syn_mcts_cfg& cfg, uint4* nodes, const uint32_t* path, uint32_t depth, uint32_t id, float v0, float v1,
                                             float v2, bool solved, bool& slow) {
    uint32_t levels = 0;
    solved = solved && cfg.solve;
    for (;;) {
        ++levels;
        uint32_t parent;
        if (!slow && !(red_exact(v0) && red_exact(v1) && red_exact(v2))) slow = true;
        if (solved || slow) {
            Rec n = load_rec(nodes, id);
            if (solved) {
                uint32_t nch = n.pk & 0xffu, nsol = (n.pk >> 8) & 0xffu;
                uint32_t cpk[9];
#pragma unroll
                for (uint32_t k = 0; k < 9u; ++k) cpk[k] = k < nch ? meta_words(nodes, n.fc + k)[MW_PK] : 0u; // all in flight together
                uint32_t bk = sol_key(nsol);
                bool all_solved = true;
#pragma unroll
                for (uint32_t k = 0; k < 9u; ++k) {
                    if (k < nch) {
                        uint32_t csol = (cpk[k] >> 8) & 0xffu;
                        uint32_t rs = csol ? sol_reversed(csol) : 0u;
                        all_solved = all_solved && rs != 0u;
                        uint32_t key = sol_key(rs);
                        bk = key > bk ? key : bk;
                    }
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
                    meta_words(nodes, id)[MW_PK] = (n.pk & 0xffff00ffu) | (best << 8);
                } else {
                    solved = false;
                }
            }
            store_stat(nodes, id, n.vis + 1.0f, n.o0 + v0, n.o1 + v1, n.o2 + v2);
            parent = n.parent;
        } else {
            red_stat(nodes, id, v0, v1, v2);
            parent = depth <= 1u ? 0u : (depth - 2u < (uint32_t)PATH_CAP ? path[(depth - 2u) * NT] : meta_words(nodes, id)[MW_PARENT]);
        }
        if (id == 0u) break;
        float tmp = v0; v0 = v2; v2 = tmp;
        id = parent;
        --depth;
    }
    return levels;
}

// The rest of visit() after Policy::eval (mcts.rs:384-397 child records, 409-423 stable softmax over
// the legal children in child order).  logits[col] is used for legal columns only.
__device__ __forceinline__ void write_children(uint4* nodes, const Pend& pd, const float (&logits)[9]) {
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
            store_rec(nodes, pd.fc + rank, 0.f, 0.f, 0.f, 0.f, __float_as_uint(__fdiv_rn(e[col], total)), pd.id, 0u,
                      (csol << 8) | ((uint32_t)col << 16));
            ++rank;
        }
    }
}

// mcts.rs:229-269 after the construction visit.  Cold path.
__device__ __noinline__ void add_root_noise(const KParams& p, uint32_t* ss, uint4* nodes) {
    const syn_mcts_cfg& cfg = p.cfg.mcts;
    if (cfg.noise_kind == SYN_NOISE_NONE) return;
    Rec r = load_rec(nodes, 0u);
    uint32_t nch = r.pk & 0xffu;
    if (nch < 2u) return;
    float w = cfg.noise_weight;
    float vals[9];
    for (int k = 0; k < 9; ++k) vals[k] = __fdiv_rn(1.0f, (float)nch);
    if (cfg.noise_kind == SYN_NOISE_DIRICHLET) {
        rng::Stream st;
        st.init(stream_seed(p, ss[SS_GI], SYN_STREAM_NOISE), ss[SS_NOISE_POS]);
        syn_dirichlet(st, cfg.noise_alpha, (int)nch, vals);
        ss[SS_NOISE_POS] = (uint32_t)st.pos;
    }
    for (uint32_t k = 0; k < nch; ++k) {
        uint32_t* mw = meta_words(nodes, r.fc + k);
        float pr = __uint_as_float(mw[MW_PRIOR]);
        pr = __fadd_rn(__fmul_rn(pr, __fsub_rn(1.0f, w)), __fmul_rn(w, vals[k]));
        mw[MW_PRIOR] = __float_as_uint(pr);
    }
}


// What the driver reads from a finished tree (mcts.rs:174-225, 273-306), by COLUMN.
__device__ __noinline__ void read_root(const uint4* nodes, uint32_t action_selection, RootOut& r) {
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
// per tree).  Cold functions take and return the thread's state BY VALUE: a Game passed by reference
// to a non-inlined function would pin the whole struct in local memory for the hot loop too.
// The root position of the tree lives in the slot record (SS_MY, SS_OP).  Returns phase | err << 8.
__device__ __forceinline__ uint64_t ss_load64(const uint32_t* ss, int w) { return *reinterpret_cast<const uint64_t*>(ss + w); }
__device__ __forceinline__ void ss_store64(uint32_t* ss, int w, uint64_t v) { *reinterpret_cast<uint64_t*>(ss + w) = v; }

struct ReadRoot2 { // the node layout of this file; tpg4.cuh passes its own reader
    __device__ __forceinline__ void operator()(const uint4* nodes, uint32_t cap, uint32_t action_selection, RootOut& r) const { read_root(nodes, action_selection, r); }
};

template <class RR>
__device__ __noinline__ int end_of_move(const KParams& p, uint32_t* ss, uint4* nodes, uint32_t nn, uint32_t e_done, RR reader) {
    const syn_rollout_cfg& cfg = p.cfg;
    struct { uint4* nodes; uint32_t nn, e_done; uint64_t my, op; } g = {nodes, nn, e_done, ss_load64(ss, SS_MY), ss_load64(ss, SS_OP)};
    RootOut r;
    reader(g.nodes, p.arena_nodes, cfg.action_selection, r);
    const uint32_t gi = ss[SS_GI], ply = ss[SS_PLY];
    atomicAdd(p.counters + CNT_NODES, (unsigned long long)g.nn);
    atomicAdd(p.counters + CNT_EXPLORES, (unsigned long long)g.e_done);
    if (p.search_mode) {
        size_t i = gi;
        for (int k = 0; k < 9; ++k) {
            if (p.s_child_visits) p.s_child_visits[i * 9 + k] = r.visits[k];
            if (p.s_child_sol) p.s_child_sol[i * 9 + k] = (uint8_t)r.child_sol[k];
        }
        if (p.s_root_q) { p.s_root_q[i * 3 + 0] = r.q0; p.s_root_q[i * 3 + 1] = r.q1; p.s_root_q[i * 3 + 2] = r.q2; }
        if (p.s_root_sol) p.s_root_sol[i] = (uint8_t)r.root_sol;
        if (p.s_best) p.s_best[i] = (uint8_t)r.best_action;
        if (p.s_nodes) p.s_nodes[i] = g.nn;
        atomicAdd(p.counters + CNT_GAMES, 1ull);
        return PH_NEED_GAME;
    }
    int err = 0;
    size_t row = (size_t)gi * 63 + ply;
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
    if (ply < cfg.random_actions_until) mode = 0;
    else if (ply < cfg.sample_actions_until && (best_sol == 0u || !cfg.stop_games_when_solved)) mode = 1;
    if (mode >= 0) {
        uint32_t ap = ss[SS_APOS];
        int a = sample_action_slow(stream_seed(p, gi, SYN_STREAM_ACTION), &ap, mode, r.legal, r.pi);
        ss[SS_APOS] = ap;
        if (a < 0 || a > 8) { err = DERR_BAD_WEIGHTS; a = r.best_action; }
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
    const uint32_t n = ply + 1u;
    ss[SS_PLY] = n;
    uint32_t fin = over ? over : (cfg.stop_games_when_solved ? solution : 0u);
    ss_store64(ss, SS_MY, g.my);
    ss_store64(ss, SS_OP, g.op);
    if (fin == 0u) return PH_NEW_TREE | (err << 8);
    // fill_state_info + store_rewards (alpha_zero.rs:296-338)
    uint32_t okind = 4u - sol_kind(fin); // solution.reversed(): the last mover's outcome
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t kind = okind;
        if (((n - 1u - k) & 1u) && kind != SYN_KIND_DRAW) kind = 4u - kind;
        size_t rr = (size_t)gi * 63 + k;
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
    p.game_len[gi] = n;
    atomicAdd(p.counters + CNT_ROWS, (unsigned long long)n);
    atomicAdd(p.counters + CNT_GAMES, 1ull);
    return PH_NEED_GAME | (err << 8);
}

// Takes the next game (or search root) from the global counter.  Cold path.
__device__ __noinline__ int next_game(const KParams& p, uint32_t* ss) {
    uint32_t gi = atomicAdd(p.next_game, 1u);
    if (gi >= p.num_games || *(volatile int*)p.error != 0) return PH_DONE;
    ss[SS_GI] = gi; ss[SS_PLY] = 0u; ss[SS_APOS] = 0u; ss[SS_FPU_POS] = 0u; ss[SS_NOISE_POS] = 0u;
    ss_store64(ss, SS_MY, p.search_mode ? p.pos_my[gi] : 0ull);
    ss_store64(ss, SS_OP, p.search_mode ? p.pos_op[gi] : 0ull);
    return PH_NEW_TREE;
}

} // namespace tp2

namespace eng {

// One persistent CTA per SM, TEAMS teams of 128 threads sharing SLOTS MLP slots (mlp_team.cuh).
// FPUK = -1: the configured syn_fpu_kind is dispatched at run time (the tuned kernels); tp2::FPU_NORMAL_CACHED: a kernel for the shipped
// Normal first-play urgency alone, with the cached stream.
template <int TEAMS, int SLOTS, bool PROF, int FPUK = -1>
__global__ void __launch_bounds__(128 * TEAMS, 1) selfplay_nn_tpg2_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ unsigned long long s_cnt[CNT_ALL];
    // the six per-round statistics are kept per WARP (plain 64-bit adds by lane 0, nobody else touches the row): shared-memory
    // 64-bit atomics on six CTA-wide counters were 3.8 % of the stall samples (profiles/r1s_*)
    __shared__ unsigned long long s_wcnt[4 * TEAMS][tp2::WC_N];
    mlpteam::Smem<TEAMS, SLOTS>& ms = *reinterpret_cast<mlpteam::Smem<TEAMS, SLOTS>*>(smem_raw);
    constexpr int NT = 128 * TEAMS, PATH_CAP = tp2::path_cap(TEAMS);
    uint32_t* const path = reinterpret_cast<uint32_t*>(smem_raw + sizeof(mlpteam::Smem<TEAMS, SLOTS>)) + threadIdx.x; // [PATH_CAP][NT] after the MLP state
    if (threadIdx.x < CNT_ALL) s_cnt[threadIdx.x] = 0ull;
    for (int i = threadIdx.x; i < 4 * TEAMS * tp2::WC_N; i += 128 * TEAMS) (&s_wcnt[0][0])[i] = 0ull;
    unsigned long long* const wc = s_wcnt[threadIdx.x >> 5];
    mlpteam::setup<TEAMS, SLOTS>(ms, p.weight_image);
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
#if SYN_TPG2_DISPATCH == 1
                // <5, ., false> never sees Fpu Normal (engine.cu::launch_tpg sends it to the FPU_NORMAL_CACHED instantiation): without the third
                // inlined descent the kernel takes 92 registers instead of 96 and runs 1.4 % faster (1,520 against 1,498 M explores/s)
                else if (cfg.fpu_kind == SYN_FPU_PARENT_Q || (TEAMS == 5 && !PROF)) err = tp2::descend<CW, SYN_FPU_PARENT_Q, NT, PATH_CAP>(p, ss, g, root, my, op, pd, rc, path);
#else
                else if (cfg.fpu_kind == SYN_FPU_PARENT_Q) err = tp2::descend<CW, SYN_FPU_PARENT_Q, NT, PATH_CAP>(p, ss, g, root, my, op, pd, rc, path);
#endif
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
        if (!mlpteam::team_any(team, g.phase != PH_DONE)) break; // no thread of this team has a game left
        uint32_t mma_phase;
        const int slot = mlpteam::acquire_slot<TEAMS, SLOTS>(ms, team, r, mma_phase);
        long long t2 = PROF ? clock64() : 0;
        if (need) mlpteam::write_features(ms.a[slot], ms.col_lut, r, my, op);
        float y[12];
#if SYN_TPG2_CBIAS
        mlpteam::forward_cb<TEAMS, SLOTS>(ms, p.mlp_bias, team, slot, r, mma_phase, y); // biases as constant-bank operands
#else
        mlpteam::forward<TEAMS, SLOTS>(ms, team, slot, r, mma_phase, y);
#endif
        mlpteam::release_slot<TEAMS, SLOTS>(ms, team, r, slot, mma_phase);
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
    mlpteam::teardown<TEAMS, SLOTS>(ms); // ends with a CTA barrier: every warp's counters are in s_cnt / s_wcnt
    __syncthreads();
    if (threadIdx.x < CNT_ALL && s_cnt[threadIdx.x]) atomicAdd(p.counters + threadIdx.x, s_cnt[threadIdx.x]);
    if (threadIdx.x < tp2::WC_N) {
        unsigned long long t = 0ull;
        for (int w = 0; w < 4 * TEAMS; ++w) t += s_wcnt[w][threadIdx.x];
        if (t) atomicAdd(p.counters + tp2::wc_counter(threadIdx.x), t);
    }
}

} // namespace eng
