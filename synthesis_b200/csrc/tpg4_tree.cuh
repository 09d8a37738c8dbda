// tpg4_tree.cuh — thread-per-game MCTS on FAMILY BLOCKS: what select_best_child reads of a node's children is ONE
// 128-byte line.
//
// Replaces synthesis/src/mcts.rs:29-489 of the reference (Node, MCTS::{explore, select_best_child, exploit_value,
// explore_value, visit, backprop, target_policy, target_q, best_action, solution}); same algorithm, same f32
// operations in the same order, same results bit for bit as tpg2.cuh / tree.cuh.
//
// Why another layout.  scripts/probe/family_read_probe.cu (profiles/r2_family_read_probe.txt): the B200 memory system
// serves 36.5 G L2-miss SECTORS per second whatever the threads per SM (128 .. 1024) and however the 32-byte sectors
// share DRAM lines.  A family of nine 32-byte records (tpg2.cuh) is nine sector misses = 4.06 G families/s; the
// kernels ran at 85-90 % of that wall (1.39 G explores/s x ~22 sector misses per explore), which is why nothing
// that did not remove sector misses ever paid.  A family whose select data is one aligned line is FOUR sector
// misses = 9.1 G families/s.
//
// Arena = 128-byte lines.  An expanded node owns a FAMILY BLOCK of three lines holding its children (mcts.rs:163-166 keeps
// them contiguous), the child of column c in slot c:
//   line 0, "select line" (32 words): what select_best_child reads of every child — four 32-byte loads
//        w0..4    num_visits of slot 0..8 as u16 (visit counts are integers: exact up to 65,535 explores)
//        w5..13   -q() of slot c, memoised by backprop (same operands, same two IEEE operations as mcts.rs:42-44, 357)
//        w14..22  action_prob of slot c
//        w23..31  link of slot c = first line of its children block:20 | num_children:4 | solution:8; ABSENT for a full column
//   lines 1-2, "stat records": what only backprop and the driver's readout touch — one 16-byte load and store per level
//        w(32+4c .. 35+4c)  { num_visits f32, sum Lose, sum Draw, sum Win } of slot c
//        w68                owner = id of the node whose children these are (the parent link, mcts.rs:31)
//   a node with ONE legal move (the ends of games, and every step of the chains auto-extend walks through, mcts.rs:404)
//   gets a one-line block with everything about its only child in the first two sectors:
//        w0 num_visits (u16) | w1 -q() | w2 action_prob | w3 link | w4..7 stat record | w8 owner
// A node id is SINGLE:1 (bit 31: the node is an only child) | line << 4 | slot; the root is the only "child" of the one-line
// block at line 0.  Children are scanned in ascending column order = the order mcts.rs:384 pushes them, so "first strict
// maximum in child order" is unchanged; which columns have a child and how many there are comes from the position the
// thread carries down the tree anyway.
//
// Every access is as wide as its data allows: a scattered 4-byte access costs the SM's load/store unit the same 32
// wavefronts per warp as a 32-byte one, and the first version of this layout (scalar loads and stores in the walk up, 127 L2
// sector requests per explore against tpg2's ~50) lost to tpg2 for that reason alone (profiles/r2_tpg4_versions.txt).
//
// Backprop (mcts.rs:429-488) reads what it updates, with exact IEEE adds in the SM (no L2 reductions: no flushing adder,
// no "slow" trees), and writes visit count, sums and memo in one pass.  Its addresses are known as soon as the descent
// ends, its values only after the leaf evaluation: prefetch_path() asks the L2 for the sums lines of the whole path
// right after the descent, the team's forward pass (tens of microseconds) hides the fetch, and the walk up then loads
// all levels together as L2 hits — one short trip for the whole way up.  While the value is "solved" the walk reads
// the node's children like the reference does, one level after the other.
//
// Compiles as device code and, under SYN_HOST_EMU (devport.cuh), as plain C++: tests/emu/tpg4_emu.cpp steps these
// functions on the CPU against the oracle, node for node.
#pragma once
#include "devport.cuh"
#include "../../include/syn_detmath.h"
#include "../../include/synthesis_b200.h"
#include "c4.cuh"
#include "treedefs.cuh"

namespace tp4 {

using namespace eng;
using tpx::Pend;
using tpx::RoundCnt;
using tpx::RootOut;
using tpx::K_NONE;
using tpx::K_LEAF;
using tpx::K_TERMINAL;
using tpx::K_INIT;

constexpr uint32_t FC_MASK = 0xfffffu;  // first line of the children block: 20 bits
constexpr uint32_t MAX_LINES = 1u << 20;
constexpr uint32_t SINGLE = 0x80000000u; // id flag: the node is an only child, in a one-line block
constexpr uint32_t ROOT_ID = SINGLE;     // slot 0 of the one-line block at line 0
constexpr uint32_t ABSENT = 0xffffffffu; // link of a slot whose column is full
constexpr uint32_t MAX_EXPLORES = 65534u; // num_visits <= explores + 1 must fit the u16 of the select line
enum { W_VIS = 0, W_NQ = 5, W_PRIOR = 14, W_LINK = 23 };

__device__ __forceinline__ uint32_t* line0(uint4* arena, uint32_t b) { return reinterpret_cast<uint32_t*>(arena + 8 * (size_t)b); }
__device__ __forceinline__ const uint32_t* line0(const uint4* arena, uint32_t b) { return reinterpret_cast<const uint32_t*>(arena + 8 * (size_t)b); }
__device__ __forceinline__ uint32_t id_line(uint32_t id) { return (id >> 4) & FC_MASK; }
__device__ __forceinline__ uint32_t id_slot(uint32_t id) { return id & 15u; }
// Word indices relative to the node's select line: its stat record {visits, Lose, Draw, Win}, its block's owner, and the
// words of the select line that hold its visit count, memo, prior and link.
__device__ __forceinline__ uint32_t stat_word(uint32_t id) { return (id & SINGLE) ? 4u : 32u + 4u * id_slot(id); }
__device__ __forceinline__ uint32_t owner_word(uint32_t id) { return (id & SINGLE) ? 8u : 68u; }
__device__ __forceinline__ uint32_t vis_word(uint32_t id) { return (id & SINGLE) ? 0u : (uint32_t)W_VIS + (id_slot(id) >> 1); }
__device__ __forceinline__ uint32_t nq_word(uint32_t id) { return (id & SINGLE) ? 1u : (uint32_t)W_NQ + id_slot(id); }
__device__ __forceinline__ uint32_t link_word(uint32_t id) { return (id & SINGLE) ? 3u : (uint32_t)W_LINK + id_slot(id); }
__device__ __forceinline__ uint32_t link_fc(uint32_t link) { return link & FC_MASK; }
__device__ __forceinline__ uint32_t link_nch(uint32_t link) { return (link >> 20) & 15u; }
__device__ __forceinline__ uint32_t link_sol(uint32_t link) { return link >> 24; }
__device__ __forceinline__ uint32_t vis_of(uint32_t word, uint32_t k) { return (word >> (16u * (k & 1u))) & 0xffffu; }

// -child.q() (mcts.rs:42-44, 357) from the sums a backprop has just produced.
__device__ __forceinline__ float memo_nq(float o0, float o2, float vis) { return -__fdiv_rn(__fsub_rn(o2, o0), vis); }

__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef SYN_HOST_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// Eight words = one 32-byte sector, one request (LDG.E.256 / STG.E.256).
__device__ __forceinline__ void load32(const uint32_t* p, uint32_t* w) {
#ifdef SYN_HOST_EMU
    for (int i = 0; i < 8; ++i) w[i] = p[i];
#else
    unsigned long long q0, q1, q2, q3;
    asm volatile("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(p) : "memory");
    w[0] = (uint32_t)q0; w[1] = (uint32_t)(q0 >> 32); w[2] = (uint32_t)q1; w[3] = (uint32_t)(q1 >> 32);
    w[4] = (uint32_t)q2; w[5] = (uint32_t)(q2 >> 32); w[6] = (uint32_t)q3; w[7] = (uint32_t)(q3 >> 32);
#endif
}
__device__ __forceinline__ void store32(uint32_t* p, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5, uint32_t w6, uint32_t w7) {
#ifdef SYN_HOST_EMU
    p[0] = w0; p[1] = w1; p[2] = w2; p[3] = w3; p[4] = w4; p[5] = w5; p[6] = w6; p[7] = w7;
#else
    const unsigned long long q0 = (unsigned long long)w0 | ((unsigned long long)w1 << 32), q1 = (unsigned long long)w2 | ((unsigned long long)w3 << 32);
    const unsigned long long q2 = (unsigned long long)w4 | ((unsigned long long)w5 << 32), q3 = (unsigned long long)w6 | ((unsigned long long)w7 << 32);
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(q0), "l"(q1), "l"(q2), "l"(q3) : "memory");
#endif
}
__device__ __forceinline__ void zero32(uint32_t* p) { store32(p, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u); }
// num_visits (u16) of a slot of a select line += 1: a fire-and-forget integer add at the L2, nobody reads the word first
__device__ __forceinline__ void red_add_u32(uint32_t* p, uint32_t v) {
#ifdef SYN_HOST_EMU
    *p += v;
#else
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

// The select line of block b: its four sectors, all in flight together.
__device__ __forceinline__ void load_select_line(const uint4* arena, uint32_t b, uint32_t (&w)[32]) {
    const uint32_t* p = line0(arena, b);
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
    for (int j = 0; j < 4; ++j) load32(p + 8 * j, w + 8 * j);
}

// A one-line block: everything about the only child (or the root) in sectors 0 and 1, written whole so that no sector is
// fetched to complete a partial write.  mcts.rs:126-131 (the root push) and mcts.rs:384-397 for a node with one legal move.
__device__ __forceinline__ void store_single_block(uint4* arena, uint32_t b, uint32_t prior, uint32_t link, uint32_t owner) {
    uint32_t* l0 = line0(arena, b);
    store32(l0, 0u, 0u, prior, link, 0u, 0u, 0u, 0u);     // visits, memo, prior, link | stat record
    store32(l0 + 8, owner, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
}
__device__ __forceinline__ void store_empty_root(uint4* arena) { store_single_block(arena, 0u, 0u, 0u, 0u); }

// What a thread keeps of the node it stands on.
struct Cur { uint32_t id; float vis, nq; uint32_t link; };

__device__ __forceinline__ Cur load_root(const uint4* arena) { // one 16-byte load
    const uint4 v = arena[0];
    Cur c;
    c.id = ROOT_ID;
    c.vis = (float)(v.x & 0xffffu);
    c.nq = __uint_as_float(v.y);
    c.link = v.w;
    return c;
}

// One explore from the (already loaded) root down to the point where the policy is needed (mcts.rs:310-325, 327-372,
// 374-406).  `my`/`op` enter as the root position and leave as the leaf's.  Returns an error code (0 = none).
// FPU = the configured syn_fpu_kind; draw() = one sample of the Fpu::Normal closure (mcts.rs:354).
// path = this thread's column of the path table (entry l-1 = the node walked at level l, stride NT).
// nb = lines in use, nn = nodes.len() of the reference's Vec (the statistic the parity tests compare).
template <int FPU, int NT, int PATH_CAP, class Draw>
__device__ __forceinline__ int descend(const syn_mcts_cfg& cfg, uint32_t arena_lines, Draw draw, uint4* arena, uint32_t& nb, uint32_t& nn, const Cur& root,
                                       uint64_t& my, uint64_t& op, Pend& pd, RoundCnt& rc, uint32_t* path) {
    Cur c = root;
    uint32_t depth = 0;
    const bool puct = cfg.exploration_kind == SYN_EXPLORATION_POLYNOMIAL_UCT;
    for (;;) {
        const uint32_t sol = link_sol(c.link), cfc = link_fc(c.link);
        if (sol) { rc.levels = depth; pd.kind = K_TERMINAL; pd.id = c.id; pd.fc = sol; pd.depth = depth; return 0; } // mcts.rs:314-316
        if (cfc == 0u) break;
        // every column with room has a child (mcts.rs:384-397); room = bit 7c set iff column c has room
        const uint64_t occ = my | op;
        const uint64_t room = (~(occ >> 6)) & c4::ROW0;
        const uint32_t nch = (uint32_t)__popcll(room);
        uint32_t b, blink, bvis;
        float bnq;
        if (nch == 1u) { // an only child: select_best_child has nothing to compare
            const uint4 v = arena[8 * (size_t)cfc]; // its one-line block: visits, memo, prior, link
            b = (uint32_t)(__ffsll((long long)room) - 1) / 7u;
            bvis = v.x & 0xffffu; bnq = __uint_as_float(v.y); blink = v.w;
            if (FPU == SYN_FPU_NORMAL && link_sol(blink) == 0u && link_fc(blink) == 0u) (void)draw(); // exploit_value still consumes its sample (mcts.rs:354)
            c.id = SINGLE | (cfc << 4);
        } else {
            uint32_t w[32];
            load_select_line(arena, cfc, w); // one line, one trip
            // ---- select_best_child (mcts.rs:327-372): first strict maximum in child order
            const float pterm = puct ? __fsqrt_rn(c.vis) : __fsqrt_rn(__fmul_rn(cfg.c, syn_logf(c.vis)));
            const float fpu_q = FPU == SYN_FPU_PARENT_Q ? -c.nq : cfg.fpu_a; // Fpu::ParentQ = parent.q() = -(-q) (mcts.rs:353)
            bool have = false;
            float bval = 0.0f;
            b = 0u; blink = 0u; bvis = 0u; bnq = 0.0f;
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
            for (uint32_t k = 0; k < 9u; ++k) {
                if ((room >> (7u * k)) & 1ull) {
                    const uint32_t link = w[W_LINK + k], vk = vis_of(w[W_VIS + (k >> 1)], k);
                    const float nqk = __uint_as_float(w[W_NQ + k]);
                    const uint32_t csol = link_sol(link);
                    // exploit_value of a proven child (mcts.rs:345-351) reads solution.reversed().value(): child Lose (kind 01) -> +1,
                    // Draw (10) -> 0, Win (11) -> -1; -inf when proven children are not to be selected.  The kind is the link's top two
                    // bits: bit 30 selects |1.0|, bit 31 is the sign (Draw becomes -0.0: q + u and every comparison treat it as +0.0)
                    const float qsol = cfg.select_solved_nodes ? __uint_as_float((((uint32_t)((int32_t)(link << 1) >> 31)) & 0x3f800000u) | (link & 0x80000000u))
                                                               : __uint_as_float(0xff800000u);
                    float q;
                    if (FPU == SYN_FPU_NORMAL) {
                        if (csol) q = qsol;
                        else if (link_fc(link) == 0u) q = draw();
                        else q = nqk;
                    } else {
                        q = link_fc(link) == 0u ? fpu_q : nqk;
                        q = csol ? qsol : q;
                    }
                    // 1 + num_visits is an integer below 2^24: the conversion is the exact sum
                    const float u = puct ? __fdiv_rn(__fmul_rn(__fmul_rn(cfg.c, __uint_as_float(w[W_PRIOR + k])), pterm), (float)(vk + 1u))
                                         : __fdiv_rn(pterm, __fsqrt_rn((float)vk));
                    const float value = __fadd_rn(q, u);
                    if (!have || value > bval) { have = true; b = k; bval = value; bvis = vk; bnq = nqk; blink = link; }
                }
            }
            c.id = (cfc << 4) | b;
        }
        rc.scanned += nch;
        c.vis = (float)bvis; c.nq = bnq; c.link = blink;
        { // Game::step (connect4.rs:221-233) into column b
            const uint64_t mover = my | c4::drop_bit(occ, (int)b);
            my = op;
            op = mover;
        }
        if (depth < (uint32_t)PATH_CAP) path[depth * NT] = c.id;
        if (++depth >= 64u) return DERR_DEPTH_OVERFLOW;
    }
    rc.levels = depth; // one select_best_child call per level walked
    // ---- visit (mcts.rs:374-406): a block for the children of the node; auto-extend through only-children
    for (;;) {
        const uint64_t occ = my | op;
        const uint64_t win = c4::winning_cells(my);   // where the mover completes four in a row
        const bool last = __popcll(occ) == 62;        // this move fills the board: a draw unless it wins
        uint32_t lm = 0u, cs2 = 0u, n = 0u;
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
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
        const uint32_t fc = nb, lines = n == 1u ? 1u : 3u;
        if (fc + lines > arena_lines) return DERR_ARENA_OVERFLOW;
        line0(arena, id_line(c.id))[link_word(c.id)] = c.link | fc | (n << 20); // mark_visited (mcts.rs:399-400)
        nb = fc + lines;
        nn += n;
        rc.expansions += 1u;
        rc.created += n;
        if (cfg.auto_extend && n == 1u) { // mcts.rs:404-405: recurse into the only child, no policy call
            const int only = __ffs((int)lm) - 1;
            const uint32_t s2 = (cs2 >> (2 * only)) & 3u;
            const uint32_t osol = s2 == 1u ? c4::SOL_LOSE0 : (s2 == 2u ? c4::SOL_DRAW0 : 0u);
            store_single_block(arena, fc, __float_as_uint(1.0f), osol << 24, c.id);
            c4::step(my, op, only);
            c.id = SINGLE | (fc << 4);
            c.vis = 0.0f; c.nq = 0.0f; c.link = osol << 24;
            if (depth < (uint32_t)PATH_CAP) path[depth * NT] = c.id;
            if (++depth >= 64u) return DERR_DEPTH_OVERFLOW;
            if (osol) { pd.kind = K_TERMINAL; pd.id = c.id; pd.fc = osol; pd.depth = depth; return 0; } // mcts.rs:377-379
            continue;
        }
        pd.kind = K_LEAF; pd.id = c.id; pd.fc = fc; pd.lc = lm | (cs2 << 9); pd.depth = depth;
        return 0;
    }
}

// Node of level l of the current path: the path table below PATH_CAP + 1, the root at 0.
template <int NT>
__device__ __forceinline__ uint32_t path_node(const uint32_t* path, uint32_t l) { return l == 0u ? ROOT_ID : path[(l - 1u) * NT]; }

// Right after the descent: ask the L2 for the stat record of every node on the path (the select lines were read on the
// way down; the stat records were not).  Nobody waits: the walk up happens after the leaf evaluation.
template <int NT, int PATH_CAP>
__device__ __forceinline__ void prefetch_path(const uint4* arena, const uint32_t* path, const Pend& pd) {
    const uint32_t top = pd.depth < (uint32_t)PATH_CAP ? pd.depth : (uint32_t)PATH_CAP;
    for (uint32_t l = 1u; l <= top; ++l) { // the root's record shares a sector with what every round reads of it
        const uint32_t n = path_node<NT>(path, l);
        if (!(n & SINGLE)) prefetch_l2(line0(arena, id_line(n)) + stat_word(n));
    }
}

// The children block of a freshly expanded node with two or more legal moves (mcts.rs:384-397): zero visits and sums, priors
// and proven outcomes by column, the owner.  Nine 32-byte stores: every sector that will ever be read is written whole.
__device__ __forceinline__ void store_block(uint4* arena, uint32_t b, const uint32_t (&prior)[9], const uint32_t (&link)[9], uint32_t owner) {
    uint32_t* l0 = line0(arena, b);
    zero32(l0);                                                                                        // visits, memo 0..2
    store32(l0 + 8, 0u, 0u, 0u, 0u, 0u, 0u, prior[0], prior[1]);                                       // memo 3..8, prior 0..1
    store32(l0 + 16, prior[2], prior[3], prior[4], prior[5], prior[6], prior[7], prior[8], link[0]);   // prior 2..8, link 0
    store32(l0 + 24, link[1], link[2], link[3], link[4], link[5], link[6], link[7], link[8]);          // link 1..8
    zero32(l0 + 32);                                                                                   // stat records 0..8
    zero32(l0 + 40);
    zero32(l0 + 48);
    zero32(l0 + 56);
    store32(l0 + 64, 0u, 0u, 0u, 0u, owner, 0u, 0u, 0u);                                               // stat record 8, owner (w68)
}

// The rest of visit() after Policy::eval (mcts.rs:409-423: stable softmax over the legal children in child order).
// logits[col] is used for legal columns only.
__device__ __forceinline__ void write_children(uint4* arena, const Pend& pd, const float (&logits)[9]) {
    const uint32_t legal = pd.lc & 0x1ffu, csol2 = pd.lc >> 9;
    float e[9];
    float total = 0.0f;
    float mx = __uint_as_float(0xff800000u);
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
    for (int col = 0; col < 9; ++col)
        if ((legal >> col) & 1u) mx = fmaxf(mx, logits[col]);
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
    for (int col = 0; col < 9; ++col) {
        e[col] = 0.0f;
        if ((legal >> col) & 1u) {
            e[col] = syn_expf(__fsub_rn(logits[col], mx));
            total = __fadd_rn(total, e[col]);
        }
    }
    uint32_t prior[9], link[9];
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
    for (int col = 0; col < 9; ++col) {
        const uint32_t s2 = (csol2 >> (2 * col)) & 3u;
        const uint32_t csol = s2 == 1u ? c4::SOL_LOSE0 : (s2 == 2u ? c4::SOL_DRAW0 : 0u);
        const bool on = (legal >> col) & 1u;
        prior[col] = on ? __float_as_uint(__fdiv_rn(e[col], total)) : 0u;
        link[col] = on ? (csol << 24) : ABSENT;
    }
    if ((legal & (legal - 1u)) == 0u) { // one legal move (auto_extend off, or the root): a one-line block
        const int only = __ffs((int)legal) - 1;
        uint32_t pr = 0u, lk = 0u;
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
        for (int col = 0; col < 9; ++col)
            if (col == only) { pr = prior[col]; lk = link[col]; }
        store_single_block(arena, pd.fc, pr, lk, pd.id);
    } else {
        store_block(arena, pd.fc, prior, link, pd.id);
    }
}

// visit() after RolloutPolicy::eval: all logits are 0, so the stable softmax is exp(0) = 1 per legal child over a total of n.
__device__ __forceinline__ void write_children_uniform(uint4* arena, const Pend& pd) {
    const uint32_t legal = pd.lc & 0x1ffu, csol2 = pd.lc >> 9;
    const uint32_t pu = __float_as_uint(__fdiv_rn(1.0f, (float)__popc(legal)));
    uint32_t prior[9], link[9];
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
    for (int col = 0; col < 9; ++col) {
        const uint32_t s2 = (csol2 >> (2 * col)) & 3u;
        const uint32_t csol = s2 == 1u ? c4::SOL_LOSE0 : (s2 == 2u ? c4::SOL_DRAW0 : 0u);
        const bool on = (legal >> col) & 1u;
        prior[col] = on ? pu : 0u;
        link[col] = on ? (csol << 24) : ABSENT;
    }
    if ((legal & (legal - 1u)) == 0u) {
        const int only = __ffs((int)legal) - 1;
        uint32_t lk = 0u;
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
        for (int col = 0; col < 9; ++col)
            if (col == only) lk = link[col];
        store_single_block(arena, pd.fc, pu, lk, pd.id);
    } else {
        store_block(arena, pd.fc, prior, link, pd.id);
    }
}

// What a walk up loads of node `id`: its stat record, one 16-byte load.
struct NodeVals { float vis, s0, s1, s2; };
__device__ __forceinline__ NodeVals load_vals(const uint4* arena, uint32_t id) {
    const uint4 v = *reinterpret_cast<const uint4*>(line0(arena, id_line(id)) + stat_word(id));
    NodeVals n;
    n.vis = __uint_as_float(v.x); n.s0 = __uint_as_float(v.y); n.s1 = __uint_as_float(v.z); n.s2 = __uint_as_float(v.w);
    return n;
}

// One level of mcts.rs:475-479 on loaded values: the stat record (one 16-byte store), the memo, and the select line's
// copy of the visit count (an integer add at the L2: the u16 shares its word with a sibling's).
__device__ __forceinline__ void bump(uint4* arena, uint32_t id, const NodeVals& n, float v0, float v1, float v2) {
    const float nv = __fadd_rn(n.vis, 1.0f);
    const float o0 = __fadd_rn(n.s0, v0), o1 = __fadd_rn(n.s1, v1), o2 = __fadd_rn(n.s2, v2);
    uint32_t* l0 = line0(arena, id_line(id));
    *reinterpret_cast<uint4*>(l0 + stat_word(id)) = make_uint4(__float_as_uint(nv), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2));
    l0[nq_word(id)] = __float_as_uint(memo_nq(o0, o2, nv));
    red_add_u32(l0 + vis_word(id), (id & SINGLE) ? 1u : 1u << (16u * (id_slot(id) & 1u)));
}

// mcts.rs:429-488 from node `id` (at level `depth`) up to the root.  While the value is "solved" the node and its children's
// solutions are read like the reference does, one level after the other.  Once it is not (the common case from the first
// level on) the rest of the way only accumulates and is known from the path table: up to CH nodes are loaded together (L2
// hits: prefetch_path asked for them before the leaf evaluation) and then updated.  Beyond PATH_CAP levels the owner link
// of the node's block is followed.
template <int NT, int PATH_CAP, int CH>
__device__ __forceinline__ void backprop(const syn_mcts_cfg& cfg, uint4* arena, const uint32_t* path, uint32_t depth, uint32_t id, float v0, float v1, float v2,
                                         bool solved) {
    solved = solved && cfg.solve;
    uint32_t lvl = depth;
    while (solved) {
        uint32_t* l0 = line0(arena, id_line(id));
        const uint32_t link = l0[link_word(id)];
        const NodeVals nv = load_vals(arena, id);
        const uint32_t owner = l0[owner_word(id)];
        const uint32_t nfc = link_fc(link), nsol = link_sol(link);
        // the children's links: one for an only child, the nine slots (ABSENT where the column is full) otherwise — the last
        // two sectors of their select line; none for a leaf
        uint32_t cl[9];
        {
            uint32_t t[16];
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
            for (int j = 0; j < 16; ++j) t[j] = ABSENT;
            if (nfc != 0u) {
                const uint32_t* c0 = line0(arena, nfc);
                if (link_nch(link) == 1u) t[7] = c0[3];
                else { load32(c0 + 16, t); load32(c0 + 24, t + 8); }
            }
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
            for (int j = 0; j < 9; ++j) cl[j] = t[7 + j]; // w23..31
        }
        uint32_t bk = sol_key(nsol);
        bool all_solved = true;
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
        for (uint32_t j = 0; j < 9u; ++j) {
            if (cl[j] != ABSENT) {
                uint32_t csol = link_sol(cl[j]);
                uint32_t rs = csol ? sol_reversed(csol) : 0u;
                all_solved = all_solved && rs != 0u;
                uint32_t key = sol_key(rs);
                bk = key > bk ? key : bk;
            }
        }
        const uint32_t best = sol_from_key(bk);
        bool mark = false;
        int slot = 0;
        if (sol_kind(best) == SYN_KIND_WIN) { mark = true; slot = 2; }
        else if (best != 0u && all_solved) { mark = true; slot = sol_kind(best) == SYN_KIND_DRAW ? 1 : 0; }
        if (mark) {
            if (cfg.correct_values_on_solve) {
                v0 = -nv.s0; v1 = -nv.s1; v2 = -nv.s2;
                const float add = __fadd_rn(nv.vis, 1.0f);
                if (slot == 2) v2 = __fadd_rn(v2, add);
                else if (slot == 1) v1 = __fadd_rn(v1, add);
                else v0 = __fadd_rn(v0, add);
            }
            l0[link_word(id)] = (link & 0x00ffffffu) | (best << 24);
        } else {
            solved = false;
        }
        bump(arena, id, nv, v0, v1, v2);
        if (lvl == 0u) return;
        float tmp = v0; v0 = v2; v2 = tmp;
        id = owner;
        --lvl;
    }
    // ---- nothing left to prove: levels lvl .. 0 only accumulate
    while (lvl > (uint32_t)PATH_CAP) { // the parent is not in the path table: follow the owner links
        const NodeVals n = load_vals(arena, id);
        const uint32_t owner = line0(arena, id_line(id))[owner_word(id)];
        bump(arena, id, n, v0, v1, v2);
        float tmp = v0; v0 = v2; v2 = tmp;
        id = owner;
        --lvl;
    }
    for (;;) { // node of level l: `id` for l == lvl, the path table below it, the root at 0
        uint32_t ids[CH];
        NodeVals nv[CH];
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
        for (uint32_t j = 0; j < (uint32_t)CH; ++j) {
            if (j <= lvl) {
                ids[j] = j == 0u ? id : path_node<NT>(path, lvl - j);
                nv[j] = load_vals(arena, ids[j]);
            }
        }
#ifndef SYN_HOST_EMU
#pragma unroll
#endif
        for (uint32_t j = 0; j < (uint32_t)CH; ++j) {
            if (j <= lvl) {
                bump(arena, ids[j], nv[j], v0, v1, v2);
                float tmp = v0; v0 = v2; v2 = tmp;
            }
        }
        if (lvl < (uint32_t)CH) break;
        lvl -= (uint32_t)CH;
        id = path_node<NT>(path, lvl);
    }
}

// What the driver reads from a finished tree (mcts.rs:174-225, 273-306), by COLUMN.  `occ` = the root position's stones.
__device__ __noinline__ void read_root(const uint4* arena, uint64_t occ, uint32_t action_selection, RootOut& r) {
    const uint32_t* r0 = line0(arena, 0u);
    const uint32_t rlink = r0[link_word(ROOT_ID)];
    const NodeVals rv = load_vals(arena, ROOT_ID);
    const float rvis = rv.vis;
    const uint32_t nch = link_nch(rlink), rsol = link_sol(rlink), rfc = link_fc(rlink);
    for (int k = 0; k < 9; ++k) { r.pi[k] = 0.0f; r.visits[k] = 0.0f; r.child_sol[k] = 0u; }
    float total = 0.0f, b0 = 0.0f, b1 = 0.0f;
    int best = 0;
    bool have = false;
    uint32_t legal = 0u;
    const uint32_t* c0 = line0(arena, rfc);
    for (uint32_t col = 0; col < 9u; ++col) {
        if ((occ >> (7u * col + 6u)) & 1ull) continue; // full column: no child
        const uint32_t cid = nch == 1u ? (SINGLE | (rfc << 4)) : ((rfc << 4) | col);
        const uint32_t link = c0[link_word(cid)];
        const NodeVals cv = load_vals(arena, cid);
        const float cvis = cv.vis;
        const uint32_t csol = link_sol(link);
        float v; // target_policy (mcts.rs:174-211)
        if (rvis == 1.0f) v = sol_kind(rsol) == SYN_KIND_WIN ? (sol_kind(csol) == SYN_KIND_LOSE ? 1.0f : 0.0f) : 1.0f;
        else v = cvis;
        total = __fadd_rn(total, v);
        float k0, k1; // best_action (mcts.rs:273-294): key (k0, k1), strict lexicographic >, first child incumbent
        const uint32_t kind = sol_kind(csol);
        if (kind == SYN_KIND_WIN) { k0 = 0.0f; k1 = (float)(csol & 63u); }
        else if (kind == 0u) { k0 = 1.0f; k1 = action_selection == SYN_ACTION_Q ? -__fdiv_rn(__fsub_rn(cv.s2, cv.s0), cvis) : cvis; }
        else if (kind == SYN_KIND_DRAW) { k0 = 2.0f; k1 = -(float)(csol & 63u); }
        else { k0 = 3.0f; k1 = -(float)(csol & 63u); }
        if (!have || k0 > b0 || (k0 == b0 && k1 > b1)) { have = true; b0 = k0; b1 = k1; best = (int)col; }
        legal |= 1u << col;
        r.pi[col] = v; r.visits[col] = cvis; r.child_sol[col] = csol;
    }
    for (int col = 0; col < 9; ++col) r.pi[col] = __fdiv_rn(r.pi[col], total); // illegal columns: 0 / total = 0
    r.legal = legal;
    r.best_action = best;
    r.root_sol = rsol;
    if (rsol) { // target_q (mcts.rs:213-225)
        int idx = sol_index(rsol);
        r.q0 = idx == 0 ? 1.0f : 0.0f; r.q1 = idx == 1 ? 1.0f : 0.0f; r.q2 = idx == 2 ? 1.0f : 0.0f;
    } else {
        r.q0 = __fdiv_rn(rv.s0, rvis); r.q1 = __fdiv_rn(rv.s1, rvis); r.q2 = __fdiv_rn(rv.s2, rvis);
    }
}

} // namespace tp4
