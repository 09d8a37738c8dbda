// tree.cuh — one MCTS tree per lane-group, device side.
//
// Replaces synthesis/src/mcts.rs:29-489 of the reference (Node, MCTS::{with_capacity, explore_n,
// explore, select_best_child, exploit_value, explore_value, visit, backprop, target_policy,
// target_q, best_action, solution, add_root_noise}).
//
// Execution model: a group of GL lanes (32 = a warp, or 16 = half a warp) owns one game and its
// tree.  Lane c works on child c of the node being scored / column c of the board being
// expanded, so the <= 9-way loops of the reference become one instruction per level.  A tree is a
// strictly serial object in the reference (one explore after another, single writer), and it
// stays one here: there is never more than one explore of a tree in flight, which is what keeps
// visit counts bit-identical; parallelism comes from the thousands of independent games.
//
// Node arena (HBM, one slab per game slot, reset every move like the reference's fresh Vec): one
// 32-byte record per node = exactly one DRAM/L2 sector,
//   bytes  0..15  stat = float4 { num_visits, outcome_probs[0] (Lose), [1] (Draw), [2] (Win) }
//   bytes 16..31  meta = uint4  { action_prob bits, first_child, parent, num_children | solution<<8 | action<<16 }
// Children of a node are contiguous (mcts.rs:163-166), so one level of selection reads <= 9
// consecutive records = <= 288 contiguous bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/syn_detmath.h"
#include "../../include/syn_sampling.h"
#include "../../include/syn_streams.h"
#include "../../include/synthesis_b200.h"
#include "c4.cuh"
#include "treedefs.cuh"
#include "rng.cuh"

namespace eng {

enum Counter {
    CNT_EXPLORES = 0, CNT_LEAF_EVALS, CNT_ROWS, CNT_GAMES, CNT_TREES, CNT_NODES, CNT_SELECT_LEVELS,
    CNT_CHILDREN_SCANNED, CNT_EXPANSIONS, CNT_CHILDREN_CREATED, CNT_BACKPROP_LEVELS, CNT_ROLLOUT_PLIES, CNT_N,
    // per-warp phase clocks of the thread-per-game kernels (lane 0 of every warp; syn_engine_debug_counters)
    DBG_T_ADVANCE = CNT_N, DBG_T_TEAMWAIT, DBG_T_MLP, DBG_T_FINISH, DBG_ROUNDS, DBG_LEAVES, DBG_T_TOTAL,
    // finer clocks of the lane-group kernels in a -DSYN_LG_PROF build: select loop, expansion, end of move, backprop
    DBG_X_SELECT, DBG_X_EXPAND, DBG_X_EOM, DBG_X_BACKPROP, CNT_ALL
};


// ------------------------------------------------------------------ lane groups
template <int GL>
struct Grp {
    unsigned mask;
    int base, gl;
    __device__ __forceinline__ Grp() {
        int lane = threadIdx.x & 31;
        base = (GL == 32) ? 0 : (lane & ~(GL - 1));
        gl = lane - base;
        mask = (GL == 32) ? 0xffffffffu : (((1u << GL) - 1u) << base);
    }
    __device__ __forceinline__ unsigned ballot(bool p) const {
        unsigned b = __ballot_sync(mask, p);
        return (GL == 32) ? b : ((b >> base) & ((1u << GL) - 1u));
    }
    __device__ __forceinline__ bool any(bool p) const { return ballot(p) != 0u; }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) const { return __shfl_sync(mask, v, base + src); }
    __device__ __forceinline__ float shfl(float v, int src) const { return __shfl_sync(mask, v, base + src); }
    __device__ __forceinline__ uint64_t shfl(uint64_t v, int src) const {
        uint32_t lo = __shfl_sync(mask, (uint32_t)v, base + src);
        uint32_t hi = __shfl_sync(mask, (uint32_t)(v >> 32), base + src);
        return ((uint64_t)hi << 32) | lo;
    }
    __device__ __forceinline__ uint32_t rmax(uint32_t v) const { return __reduce_max_sync(mask, v); }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};

__device__ __forceinline__ uint32_t float_sort_key(float f) { // monotone for non-NaN floats, -0 == +0
    uint32_t u = __float_as_uint(f);
    if (u == 0x80000000u) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__constant__ uint32_t GEN_RANGE_ZONE[10] = { // rand 0.8 UniformInt<u8> zone for n = 0..9 (u32-widened)
    0u,
    0xffffffffu - ((0xffffffffu - 1u + 1u) % 1u), 0xffffffffu - ((0xffffffffu - 2u + 1u) % 2u),
    0xffffffffu - ((0xffffffffu - 3u + 1u) % 3u), 0xffffffffu - ((0xffffffffu - 4u + 1u) % 4u),
    0xffffffffu - ((0xffffffffu - 5u + 1u) % 5u), 0xffffffffu - ((0xffffffffu - 6u + 1u) % 6u),
    0xffffffffu - ((0xffffffffu - 7u + 1u) % 7u), 0xffffffffu - ((0xffffffffu - 8u + 1u) % 8u),
    0xffffffffu - ((0xffffffffu - 9u + 1u) % 9u)};

// ------------------------------------------------------------------ rollout stream (hot)
// The group regenerates 4*GL words (GL/4 ChaCha12 blocks) at a time into shared memory: four
// lanes per block, lane j of a quad holding column j of the 4x4 state; the diagonal rounds rotate
// rows 1..3 across the quad with shuffles.
template <int GL>
struct RolloutRng {
    static constexpr int WORDS = 4 * GL;
    uint32_t* buf; // shared, WORDS entries, private to the group
    uint32_t kA, kB; // this lane's two key words: key[j], key[4+j], j = gl & 3
    uint32_t pos;    // words drawn so far
    __device__ __forceinline__ void init(const Grp<GL>& g, uint64_t seed, uint32_t* smem) {
        uint32_t key[8];
        rng::seed_key(seed, key);
        int j = g.gl & 3;
        kA = j == 0 ? key[0] : (j == 1 ? key[1] : (j == 2 ? key[2] : key[3]));
        kB = j == 0 ? key[4] : (j == 1 ? key[5] : (j == 2 ? key[6] : key[7]));
        buf = smem;
        pos = 0;
    }
    __device__ __noinline__ void refill(const Grp<GL>& g) {
        g.sync();
        const int j = g.gl & 3, q = g.gl & ~3;
        const uint64_t ctr = (uint64_t)(pos >> 4) + (uint64_t)(g.gl >> 2);
        const uint32_t c0 = j == 0 ? 0x61707865u : (j == 1 ? 0x3320646eu : (j == 2 ? 0x79622d32u : 0x6b206574u));
        const uint32_t c3 = j == 0 ? (uint32_t)ctr : (j == 1 ? (uint32_t)(ctr >> 32) : 0u);
        uint32_t a = c0, b = kA, c = kB, d = c3;
#pragma unroll 1
        for (int r = 0; r < 6; ++r) {
            SYN_QR(a, b, c, d)
            b = g.shfl(b, q + ((j + 1) & 3));
            c = g.shfl(c, q + ((j + 2) & 3));
            d = g.shfl(d, q + ((j + 3) & 3));
            SYN_QR(a, b, c, d)
            b = g.shfl(b, q + ((j + 3) & 3));
            c = g.shfl(c, q + ((j + 2) & 3));
            d = g.shfl(d, q + ((j + 1) & 3));
        }
        uint32_t* o = buf + (g.gl >> 2) * 16 + j;
        o[0] = a + c0; o[4] = b + kA; o[8] = c + kB; o[12] = d + c3;
        g.sync();
    }
    __device__ __forceinline__ uint32_t next(const Grp<GL>& g) {
        if ((pos & (WORDS - 1)) == 0u) refill(g);
        return buf[(pos++) & (WORDS - 1)];
    }
};

// policies/rollout.rs:8-31 from a non-terminal leaf.  Returns the one-hot index (0 Lose, 1 Draw, 2 Win) for the leaf's
// player to move; uniform across the group.
//
// A playout is a chain of dependent plies (~150 dependent instructions each for a lone warp), but the chain is only as long
// as the SET of legal columns changes: while no column fills up, ply i's column is a function of stream word i alone
// (gen_range(0..n) with the same n and the same legal set).  So the group plays a WINDOW of plies at once, lane i taking
// word i of the stream: columns in parallel, the row of every stone from a ballot per column (stones before it in the
// same column), each mover's board by a prefix-OR over the lanes of the same parity, won() for every ply in parallel.
// The window ends after the first ply that fills a column (the legal set changes), before a rejected word (2^32 mod n
// zone, rand 0.8 UniformInt), at the end of the buffered words, or at the first terminal ply — whichever comes first;
// the plies before that point are exactly the sequential playout's, and so is the number of words consumed.
template <int GL>
__device__ __forceinline__ int rollout(const Grp<GL>& g, uint64_t my, uint64_t op, RolloutRng<GL>& rr, uint32_t& plies) {
    uint32_t k = 0; // plies played so far
    for (;;) {
        if ((rr.pos & (RolloutRng<GL>::WORDS - 1)) == 0u) rr.refill(g);
        const uint64_t occ = my | op;
        const uint64_t legal = (~(occ >> 6)) & c4::ROW0; // bit 7c set <=> column c has room
        const uint32_t n = (uint32_t)__popcll(legal);
        const uint32_t zone = GEN_RANGE_ZONE[n];
        const uint32_t in_buf = RolloutRng<GL>::WORDS - (rr.pos & (RolloutRng<GL>::WORDS - 1));
        const uint32_t w = in_buf < (uint32_t)GL ? in_buf : (uint32_t)GL; // lanes that hold a word
        const bool has = (uint32_t)g.gl < w;
        const uint32_t v = has ? rr.buf[(rr.pos + (uint32_t)g.gl) & (RolloutRng<GL>::WORDS - 1)] : 0u;
        const uint64_t m = (uint64_t)v * (uint64_t)n;
        const uint32_t hi = (uint32_t)(m >> 32);
        const bool rejected = has && (uint32_t)m > zone;
        const uint32_t rejm = g.ballot(rejected);
        // the hi-th legal column, ascending
        int col = 0;
        {
            uint32_t seen = 0;
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                const uint32_t bit = (uint32_t)(legal >> (7 * c)) & 1u;
                if (bit && seen == hi) col = c;
                seen += bit;
            }
        }
        // stones of this window that land in my column before mine
        uint32_t below = 0;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            const uint32_t b = g.ballot(has && col == c);
            if (col == c) below = b;
        }
        below = (uint32_t)__popc(below & ((1u << g.gl) - 1u));
        const uint32_t row = (uint32_t)__popc((uint32_t)(occ >> (7 * col)) & 0x7fu) + below;
        const bool overflow = has && row >= 7u;           // only possible after a ply of this window filled the column
        const bool fills = has && row == 6u;              // the legal set changes after this ply
        const uint32_t bad = g.ballot(rejected || overflow || !has);
        const uint32_t first_bad = bad ? (uint32_t)__ffs(bad) - 1u : (uint32_t)GL;
        const uint32_t fillm = g.ballot(fills) & ((first_bad >= 32u) ? 0xffffffffu : ((1u << first_bad) - 1u));
        uint32_t end = first_bad;                          // plies 0..end-1 of the window are the sequential playout's
        bool filled_last = false;                          // the window's last ply filled a column
        if (fillm) { const uint32_t f = (uint32_t)__ffs(fillm); if (f <= end) { end = f; filled_last = true; } }
        // every mover's board after its ply: the mover's board at the start of the window | its stones so far
        uint64_t x = (has && row < 7u) ? (1ull << (7 * col + (int)row)) : 0ull;
#pragma unroll
        for (int o = 2; o < GL; o <<= 1) {
            const uint32_t lo = __shfl_up_sync(g.mask, (uint32_t)x, o, GL), hi32 = __shfl_up_sync(g.mask, (uint32_t)(x >> 32), o, GL);
            if (g.gl >= o) x |= ((uint64_t)hi32 << 32) | lo;
        }
        x |= (g.gl & 1) ? op : my;
        const bool in_win = (uint32_t)g.gl < end;
        const bool win = in_win && c4::won(x);
        const bool full = in_win && (uint32_t)__popcll(occ) + (uint32_t)g.gl + 1u == 63u;
        const uint32_t term = g.ballot(win || full);
        if (term) {
            const int tl = __ffs(term) - 1;
            const bool was_win = (g.ballot(win) >> tl) & 1u;
            k += (uint32_t)tl + 1u;
            rr.pos += (uint32_t)tl + 1u;
            plies += k;
            return was_win ? ((k & 1u) ? 2 : 0) : 1;
        }
        // no terminal ply in the window: take all of it (and the rejected word behind it, if that is what ended it)
        if (end > 0u) {
            const uint64_t last = g.shfl(x, (int)end - 1);                               // the board of the window's last mover
            const uint64_t prev = end >= 2u ? g.shfl(x, (int)end - 2) : op;              // the other player's board
            my = prev;
            op = last;
        }
        k += end;
        // a rejected word right behind the window is consumed with it — unless the window ended on a filled column: then the next
        // word meets a different n, and whether IT is rejected is for the next window to say (tests/test_windowed_playout_logic.py)
        rr.pos += end + ((!filled_last && end == first_bad && end < w && ((rejm >> end) & 1u)) ? 1u : 0u);
    }
}

// Views of the two halves of the 32-byte node records (see the header comment).
struct StatView {
    uint4* base;
    __device__ __forceinline__ float4& operator[](size_t i) const { return *reinterpret_cast<float4*>(base + 2 * i); }
};
struct MetaView {
    uint4* base;
    __device__ __forceinline__ uint4& operator[](size_t i) const { return base[2 * i + 1]; }
};

// ------------------------------------------------------------------ per-group tree context
template <int GL>
struct Tree {
    StatView stat;  // this slot's arena
    MetaView meta;
    uint32_t* path; // shared, 64 entries: node ids root..current
    uint32_t nn;    // nodes.len()
    uint32_t cap;   // arena capacity in nodes
    const syn_mcts_cfg* cfg; // in shared/constant memory
    uint32_t cnt[CNT_N]; // flushed to the global 64-bit counters at the end of every game
    int err;
    uint32_t fpu_pos, noise_pos; // words drawn from the per-game FPU / noise streams
    uint64_t fpu_seed, noise_seed;
    uint32_t* fpu_sm = nullptr;  // shared, FPU_SM_WORDS per group: the FPU stream's key and current block (fpu_stream_begin), or null
#ifdef SYN_LG_PROF
    long long pt[4]; // cycles in {select loop, expansion, end of move, backprop}
#endif
};
#ifdef SYN_LG_PROF
#define LGP_NOW() clock64()
#define LGP_ADD(t, i, since) ((t).pt[i] += clock64() - (since))
#else
#define LGP_NOW() 0ll
#define LGP_ADD(t, i, since) ((void)(since))
#endif

struct Pending { // an expanded leaf waiting for Policy::eval
    uint32_t leaf, first_child, legal; // legal = 9-bit column mask of the leaf's children
    int depth;
    bool any_solved;
    uint64_t my, op; // the leaf's board
};

// mcts.rs:429-488.  Levels d..0 of `path`; `v` is the vector for level d.
template <int GL>
__device__ __forceinline__ void backprop(const Grp<GL>& g, Tree<GL>& t, int d, float v0, float v1, float v2, bool solved) {
    const syn_mcts_cfg& cfg = *t.cfg;
    const long long lgp_bp = LGP_NOW();
    g.sync();
    t.cnt[CNT_BACKPROP_LEVELS] += (uint32_t)(d + 1);
    // serial part: only while the solver may still mark nodes
    while (cfg.solve && solved) {
        uint32_t id = t.path[d];
        uint4 m = t.meta[id];
        float4 s = t.stat[id];
        uint32_t nch = m.w & 0xffu, nsol = (m.w >> 8) & 0xffu;
        uint32_t csol = 0u;
        bool act = (uint32_t)g.gl < nch;
        if (act) csol = (t.meta[m.y + g.gl].w >> 8) & 0xffu;
        uint32_t rs = csol ? sol_reversed(csol) : 0u;
        bool all_solved = !g.any(act && rs == 0u);
        uint32_t key = act ? sol_key(rs) : 0u;
        uint32_t bk = g.rmax(key);
        uint32_t nk = sol_key(nsol);
        if (nk > bk) bk = nk;
        uint32_t best = sol_from_key(bk);
        bool mark = false;
        int slot = 0;
        if (sol_kind(best) == SYN_KIND_WIN) { mark = true; slot = 2; }
        else if (best != 0u && all_solved) { mark = true; slot = sol_kind(best) == SYN_KIND_DRAW ? 1 : 0; }
        if (mark) {
            m.w = (m.w & 0xffff00ffu) | (best << 8);
            if (cfg.correct_values_on_solve) {
                v0 = -s.y; v1 = -s.z; v2 = -s.w;
                float add = s.x + 1.0f;
                if (slot == 2) v2 = v2 + add;
                else if (slot == 1) v1 = v1 + add;
                else v0 = v0 + add;
            }
            if (g.gl == 0) t.meta[id].w = m.w;
        } else {
            solved = false;
        }
        s.y += v0; s.z += v1; s.w += v2; s.x += 1.0f;
        if (g.gl == 0) t.stat[id] = s;
        if (d == 0) { g.sync(); LGP_ADD(t, 3, lgp_bp); return; }
        float tmp = v0; v0 = v2; v2 = tmp;
        --d;
        g.sync(); // the parent's scan must see this node's new solution
    }
    // parallel part: no more marking, every remaining ancestor just accumulates
    for (int j = g.gl; j <= d; j += GL) {
        uint32_t id = t.path[j];
        float4 s = t.stat[id];
        bool sw = ((d - j) & 1) != 0;
        s.y += sw ? v2 : v0;
        s.z += v1;
        s.w += sw ? v0 : v2;
        s.x += 1.0f;
        t.stat[id] = s;
    }
    g.sync();
    LGP_ADD(t, 3, lgp_bp);
}

// Normal FPU for the unvisited children of one parent, in child order (mcts.rs:351-355 with
// Fpu::Func = the shipped Normal(mean, std) closure).  The closure runs for EVERY unvisited child EVERY time its parent is
// selected through, so with the shipped configuration this is on the hot path of most levels: the stream's expanded key
// and its current ChaCha12 block stay in shared memory between calls (re-deriving both per call — ~800 instructions of one
// lane — made a 1,600-explore move 2.4 x slower than with a constant FPU).
constexpr int FPU_SM_WORDS = 28; // [0, 8) key, [8, 24) block, [24] index of the block held (~0u: none)
// rng::chacha12_block for a key and a destination in shared memory.  A separate function on purpose: rng::chacha12_block is
// shared (noinline) with the thread-per-game kernels, and giving it a new kind of caller changed ITS register footprint and
// with it the register allocation of selfplay_nn_tpg2_kernel around its calls: 2.5 % of the headline bench, found by
// diffing SASS between builds.
__device__ __noinline__ void chacha12_block_smem(uint32_t* sm, uint32_t counter) {
    const uint32_t k0 = sm[0], k1 = sm[1], k2 = sm[2], k3 = sm[3], k4 = sm[4], k5 = sm[5], k6 = sm[6], k7 = sm[7];
    uint32_t x0 = 0x61707865u, x1 = 0x3320646eu, x2 = 0x79622d32u, x3 = 0x6b206574u;
    uint32_t x4 = k0, x5 = k1, x6 = k2, x7 = k3, x8 = k4, x9 = k5, x10 = k6, x11 = k7;
    uint32_t x12 = counter, x13 = 0u, x14 = 0u, x15 = 0u;
#pragma unroll 1
    for (int r = 0; r < 6; ++r) {
        SYN_QR(x0, x4, x8, x12) SYN_QR(x1, x5, x9, x13) SYN_QR(x2, x6, x10, x14) SYN_QR(x3, x7, x11, x15)
        SYN_QR(x0, x5, x10, x15) SYN_QR(x1, x6, x11, x12) SYN_QR(x2, x7, x8, x13) SYN_QR(x3, x4, x9, x14)
    }
    uint32_t* out = sm + 8;
    out[0] = x0 + 0x61707865u; out[1] = x1 + 0x3320646eu; out[2] = x2 + 0x79622d32u; out[3] = x3 + 0x6b206574u;
    out[4] = x4 + k0; out[5] = x5 + k1; out[6] = x6 + k2; out[7] = x7 + k3;
    out[8] = x8 + k4; out[9] = x9 + k5; out[10] = x10 + k6; out[11] = x11 + k7;
    out[12] = x12 + counter; out[13] = x13; out[14] = x14; out[15] = x15;
}
struct FpuSmemStream { // R-concept of include/syn_sampling.h; used by lane 0 of a group
    uint32_t* sm;
    uint32_t pos;
    __device__ uint32_t next_u32() {
        const uint32_t blk = pos >> 4;
        if (blk != sm[24]) {
            chacha12_block_smem(sm, blk);
            sm[24] = blk;
        }
        return sm[8 + ((pos++) & 15u)];
    }
};
// A new game: the stream starts over under the game's FPU seed.
template <int GL>
__device__ __forceinline__ void fpu_stream_begin(const Grp<GL>& g, Tree<GL>& t) {
    if (t.fpu_sm == nullptr) return;
    g.sync();
    if (g.gl == 0) {
        uint32_t key[8];
        rng::seed_key(t.fpu_seed, key);
#pragma unroll
        for (int i = 0; i < 8; ++i) t.fpu_sm[i] = key[i];
        t.fpu_sm[24] = 0xffffffffu;
    }
    g.sync();
}
template <int GL>
__device__ __noinline__ float fpu_normal_slow(const Grp<GL>& g, Tree<GL>& t, unsigned need_mask) {
    float mine = 0.0f;
    float vals[9];
    uint32_t newpos = 0;
    if (g.gl == 0) {
        if (t.fpu_sm != nullptr) {
            FpuSmemStream st{t.fpu_sm, t.fpu_pos};
            for (int k = 0; k < 9; ++k) vals[k] = ((need_mask >> k) & 1u) ? syn_normal(st, t.cfg->fpu_a, t.cfg->fpu_b) : 0.0f;
            newpos = st.pos;
        } else {
            rng::Stream st;
            st.init(t.fpu_seed, t.fpu_pos);
            for (int k = 0; k < 9; ++k) vals[k] = ((need_mask >> k) & 1u) ? syn_normal(st, t.cfg->fpu_a, t.cfg->fpu_b) : 0.0f;
            newpos = (uint32_t)st.pos;
        }
    }
    t.fpu_pos = g.shfl(newpos, 0);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        float v = g.shfl(g.gl == 0 ? vals[k] : 0.0f, 0);
        if (g.gl == k) mine = v;
    }
    return mine;
}

// One explore up to the point where the policy is needed (mcts.rs:310-325 + 327-341 + 374-406).
// Returns true and fills `p` if a leaf awaits evaluation; false if the explore is complete
// (it ended on a solved node) or an error was raised.
template <int GL>
__device__ __forceinline__ bool explore_descend(const Grp<GL>& g, Tree<GL>& t, uint64_t my, uint64_t op, Pending& p) {
    const syn_mcts_cfg& cfg = *t.cfg;
    int d = 0;
    uint32_t cur = 0u;
    uint4 m = t.meta[0];
    float4 s = t.stat[0];
    uint32_t cfc = m.y, cpk = m.w;
    float cvis = s.x, cop0 = s.y, cop2 = s.w;
    if (g.gl == 0) t.path[0] = 0u;
    const long long lgp_sel = LGP_NOW();
    for (;;) {
        uint32_t sol = (cpk >> 8) & 0xffu, nch = cpk & 0xffu;
        if (sol) { // mcts.rs:314-316
            int idx = sol_index(sol);
            LGP_ADD(t, 0, lgp_sel);
            backprop(g, t, d, idx == 0 ? 1.0f : 0.0f, idx == 1 ? 1.0f : 0.0f, idx == 2 ? 1.0f : 0.0f, true);
            return false;
        }
        if (nch == 0u) break; // unvisited: expand below
        // ---- select_best_child (mcts.rs:327-372)
        bool act = (uint32_t)g.gl < nch;
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 cm = make_uint4(0u, 0u, 0u, 0u);
        if (act) {
            cs = t.stat[cfc + g.gl];
            cm = t.meta[cfc + g.gl];
        }
        uint32_t csol = (cm.w >> 8) & 0xffu, cn = cm.w & 0xffu;
        // (every lane prefetching ITS child's family into L2 before the winner is known was measured: select +1-10 % slower —
        // at these batch sizes the descent waits for its own instruction chain, not for DRAM; profiles/r2_lane_group_clocks.txt)
        bool unvisited = act && csol == 0u && cn == 0u;
        float q;
        if (csol) {
            uint32_t k = sol_kind(csol);
            q = cfg.select_solved_nodes ? (k == SYN_KIND_WIN ? -1.0f : (k == SYN_KIND_LOSE ? 1.0f : 0.0f)) : __uint_as_float(0xff800000u);
        } else if (cn == 0u) {
            q = cfg.fpu_kind == SYN_FPU_CONST ? cfg.fpu_a : __fdiv_rn(__fsub_rn(cop2, cop0), cvis);
        } else {
            q = -__fdiv_rn(__fsub_rn(cs.w, cs.y), cs.x);
        }
        if (cfg.fpu_kind == SYN_FPU_NORMAL) {
            unsigned need = g.ballot(unvisited);
            if (need) {
                float f = fpu_normal_slow(g, t, need);
                if (unvisited) q = f;
            }
        }
        float u;
        if (cfg.exploration_kind == SYN_EXPLORATION_POLYNOMIAL_UCT) {
            float visits = __fsqrt_rn(cvis);
            u = __fdiv_rn(__fmul_rn(__fmul_rn(cfg.c, __uint_as_float(cm.x)), visits), __fadd_rn(1.0f, cs.x));
        } else {
            float visits = __fsqrt_rn(__fmul_rn(cfg.c, syn_logf(cvis)));
            u = __fdiv_rn(visits, __fsqrt_rn(cs.x));
        }
        float value = __fadd_rn(q, u);
        // first strict maximum in child order; child 0 is the incumbent even if NaN (mcts.rs:334-338)
        float v0 = g.shfl(value, 0);
        int b = 0;
        if (v0 == v0) {
            uint32_t key = act ? float_sort_key(value == value ? value : __uint_as_float(0xff800000u)) : 0u;
            uint32_t mx = g.rmax(key);
            b = __ffs(g.ballot(act && key == mx)) - 1;
        }
        t.cnt[CNT_SELECT_LEVELS] += 1u;
        t.cnt[CNT_CHILDREN_SCANNED] += nch;
        cur = cfc + (uint32_t)b;
        cvis = g.shfl(cs.x, b);
        cop0 = g.shfl(cs.y, b);
        cop2 = g.shfl(cs.w, b);
        uint32_t nfc = g.shfl(cm.y, b);
        cpk = g.shfl(cm.w, b);
        cfc = nfc;
        c4::step(my, op, (int)((cpk >> 16) & 0xffu));
        ++d;
        if (d >= 64) { t.err = DERR_DEPTH_OVERFLOW; return false; }
        if (g.gl == 0) t.path[d] = cur;
    }
    // ---- visit (mcts.rs:374-406): push the children of `cur`; auto-extend through only-children
    LGP_ADD(t, 0, lgp_sel);
    const long long lgp_exp = LGP_NOW();
    for (;;) {
        uint64_t occ = my | op;
        int col = g.gl;
        bool legal = false;
        uint32_t csol = 0u;
        if (col < 9) {
            uint32_t colbits = (uint32_t)((occ >> (7 * col)) & 0x7full);
            legal = colbits != 0x7fu;
            if (legal) {
                uint64_t bit = 1ull << (7 * col + __popc(colbits));
                uint64_t mover = my | bit;
                if (c4::won(mover)) csol = c4::SOL_LOSE0;            // Outcome::from(reward(player)) = Lose(0)
                else if ((occ | bit) == c4::ALL) csol = c4::SOL_DRAW0; // board full: Draw(0)
            }
        }
        uint32_t lm = g.ballot(legal);
        uint32_t n = (uint32_t)__popc(lm);
        uint32_t fc = t.nn;
        if (fc + n > t.cap) { t.err = DERR_ARENA_OVERFLOW; return false; }
        if (legal) {
            uint32_t rank = (uint32_t)__popc(lm & ((1u << col) - 1u));
            t.meta[fc + rank] = make_uint4(__float_as_uint(1.0f), 0u, cur, (csol << 8) | ((uint32_t)col << 16));
            t.stat[fc + rank] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        bool any_solved = g.any(legal && csol != 0u);
        if (g.gl == 0) { // mark_visited (mcts.rs:399-400)
            t.meta[cur].y = fc;
            t.meta[cur].w = (cpk & 0xffffff00u) | n;
        }
        t.nn = fc + n;
        t.cnt[CNT_EXPANSIONS] += 1u;
        t.cnt[CNT_CHILDREN_CREATED] += n;
        if (cfg.auto_extend && n == 1u) { // mcts.rs:404-405
            int only = __ffs(lm) - 1;
            uint32_t osol = g.shfl(csol, only);
            c4::step(my, op, only);
            cur = fc;
            cpk = (osol << 8) | ((uint32_t)only << 16);
            ++d;
            if (d >= 64) { t.err = DERR_DEPTH_OVERFLOW; return false; }
            if (g.gl == 0) t.path[d] = cur;
            if (osol) { // visit() of a solved node returns its one-hot (mcts.rs:377-379)
                int idx = sol_index(osol);
                LGP_ADD(t, 1, lgp_exp);
                backprop(g, t, d, idx == 0 ? 1.0f : 0.0f, idx == 1 ? 1.0f : 0.0f, idx == 2 ? 1.0f : 0.0f, true);
                return false;
            }
            continue;
        }
        p.leaf = cur; p.first_child = fc; p.legal = lm; p.depth = d; p.any_solved = any_solved; p.my = my; p.op = op;
        LGP_ADD(t, 1, lgp_exp);
        return true;
    }
}

// The rest of visit() after Policy::eval: stable softmax over the legal children in child order
// (mcts.rs:409-423), then backprop.  `logit` is this lane's column's logit.
template <int GL>
__device__ __forceinline__ void explore_finish(const Grp<GL>& g, Tree<GL>& t, const Pending& p, bool uniform_prior, float logit,
                                               float o0, float o1, float o2) {
    int col = g.gl;
    bool legal = col < 9 && ((p.legal >> col) & 1u);
    uint32_t n = (uint32_t)__popc(p.legal);
    float prior;
    if (uniform_prior) { // all logits equal: exp(0) = 1 for every child, total = n exactly
        prior = __fdiv_rn(1.0f, (float)n);
    } else {
        float l = legal ? logit : __uint_as_float(0xff800000u);
        float mx = l;
#pragma unroll
        for (int o = GL / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(g.mask, mx, o));
        float e = legal ? syn_expf(__fsub_rn(l, mx)) : 0.0f;
        float total = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float ek = g.shfl(e, k);
            if ((p.legal >> k) & 1u) total = __fadd_rn(total, ek);
        }
        prior = __fdiv_rn(e, total);
    }
    if (legal) {
        uint32_t rank = (uint32_t)__popc(p.legal & ((1u << col) - 1u));
        t.meta[p.first_child + rank].x = __float_as_uint(prior);
    }
    t.cnt[CNT_LEAF_EVALS] += 1u;
    backprop(g, t, p.depth, o0, o1, o2, p.any_solved);
}

// mcts.rs:229-269 after the construction visit.  Cold path.
template <int GL>
__device__ __noinline__ void add_root_noise(const Grp<GL>& g, Tree<GL>& t) {
    const syn_mcts_cfg& cfg = *t.cfg;
    if (cfg.noise_kind == SYN_NOISE_NONE) return;
    g.sync();
    uint4 m = t.meta[0];
    uint32_t nch = m.w & 0xffu;
    if (nch < 2u) return;
    float w = cfg.noise_weight;
    float noise = __fdiv_rn(1.0f, (float)nch);
    if (cfg.noise_kind == SYN_NOISE_DIRICHLET) {
        float vals[9];
        uint32_t newpos = 0;
        if (g.gl == 0) {
            rng::Stream st;
            st.init(t.noise_seed, t.noise_pos);
            syn_dirichlet(st, cfg.noise_alpha, (int)nch, vals);
            newpos = (uint32_t)st.pos;
        }
        t.noise_pos = g.shfl(newpos, 0);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float v = g.shfl((g.gl == 0 && k < (int)nch) ? vals[k] : 0.0f, 0);
            if (g.gl == k) noise = v;
        }
    }
    if ((uint32_t)g.gl < nch) {
        float pr = __uint_as_float(t.meta[m.y + g.gl].x);
        pr = __fadd_rn(__fmul_rn(pr, __fsub_rn(1.0f, w)), __fmul_rn(w, noise));
        t.meta[m.y + g.gl].x = __float_as_uint(pr);
    }
    g.sync();
}

// What the driver reads from a finished tree (mcts.rs:174-225, 273-306).  Per-lane outputs are
// indexed by COLUMN (lane c = column c): pi, child visits, child solution.
struct RootReadout {
    float pi;          // target_policy[col]
    float visits;      // child num_visits (0 if illegal)
    uint32_t child_sol; // packed solution of the child playing this column
    float q0, q1, q2;  // target_q
    uint32_t root_sol;
    int best_action;
    uint32_t legal; // 9-bit mask of root children columns
};

template <int GL>
__device__ __forceinline__ void read_root(const Grp<GL>& g, Tree<GL>& t, uint32_t action_selection, RootReadout& r) {
    g.sync();
    uint4 m = t.meta[0];
    float4 s = t.stat[0];
    uint32_t nch = m.w & 0xffu, rsol = (m.w >> 8) & 0xffu;
    bool act = (uint32_t)g.gl < nch;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 cm = make_uint4(0u, 0u, 0u, 0u);
    if (act) {
        cs = t.stat[m.y + g.gl];
        cm = t.meta[m.y + g.gl];
    }
    uint32_t csol = (cm.w >> 8) & 0xffu, cact = (cm.w >> 16) & 0xffu;
    // target_policy (mcts.rs:174-211)
    float v;
    if (s.x == 1.0f) {
        if (sol_kind(rsol) == SYN_KIND_WIN) v = (sol_kind(csol) == SYN_KIND_LOSE) ? 1.0f : 0.0f;
        else v = 1.0f;
    } else {
        v = cs.x;
    }
    if (!act) v = 0.0f;
    float total = 0.0f;
    // best_action (mcts.rs:273-294): key (k0, k1), strict lexicographic >, first child incumbent
    float k0, k1;
    uint32_t kind = sol_kind(csol);
    if (kind == SYN_KIND_WIN) { k0 = 0.0f; k1 = (float)(csol & 63u); }
    else if (kind == 0u) { k0 = 1.0f; k1 = action_selection == SYN_ACTION_Q ? -__fdiv_rn(__fsub_rn(cs.w, cs.y), cs.x) : cs.x; }
    else if (kind == SYN_KIND_DRAW) { k0 = 2.0f; k1 = -(float)(csol & 63u); }
    else { k0 = 3.0f; k1 = -(float)(csol & 63u); }
    float b0 = 0.f, b1 = 0.f;
    int best = 0;
    float pi_col = 0.0f, vis_col = 0.0f;
    uint32_t sol_col = 0u, legal = 0u;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        float vk = g.shfl(v, k);
        float ck0 = g.shfl(k0, k), ck1 = g.shfl(k1, k);
        float cvis = g.shfl(cs.x, k);
        uint32_t ca = g.shfl(cact, k), csl = g.shfl(csol, k);
        if ((uint32_t)k < nch) {
            total = __fadd_rn(total, vk);
            if (k == 0 || ck0 > b0 || (ck0 == b0 && ck1 > b1)) { b0 = ck0; b1 = ck1; best = (int)ca; }
            legal |= 1u << ca;
            if ((int)ca == g.gl) { pi_col = vk; vis_col = cvis; sol_col = csl; }
        }
    }
    r.pi = __fdiv_rn(pi_col, total); // illegal columns: 0 / total = 0 (mcts.rs:208-210)
    r.visits = vis_col;
    r.child_sol = sol_col;
    r.legal = legal;
    r.best_action = best;
    r.root_sol = rsol;
    if (rsol) { // target_q (mcts.rs:213-225)
        int idx = sol_index(rsol);
        r.q0 = idx == 0 ? 1.0f : 0.0f; r.q1 = idx == 1 ? 1.0f : 0.0f; r.q2 = idx == 2 ? 1.0f : 0.0f;
    } else {
        r.q0 = __fdiv_rn(s.y, s.x); r.q1 = __fdiv_rn(s.z, s.x); r.q2 = __fdiv_rn(s.w, s.x);
    }
}

} // namespace eng
