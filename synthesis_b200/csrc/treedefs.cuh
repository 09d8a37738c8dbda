// treedefs.cuh — small definitions shared by every tree schedule (tree.cuh, tpg.cuh, tpg2.cuh, tpg4_tree.cuh):
// the packed Option<Outcome> of game.rs:9-66 and what a descent hands to the finish of an explore.
// Compiles as device code and, under SYN_HOST_EMU (devport.cuh), as plain C++ for tests/emu.
#pragma once
#include "devport.cuh"
#include "../../include/synthesis_b200.h"

namespace eng {

enum DeviceError { DERR_NONE = 0, DERR_ARENA_OVERFLOW = 1, DERR_DEPTH_OVERFLOW = 2, DERR_BAD_WEIGHTS = 3, DERR_NO_BEST_ACTION = 4 };

// ------------------------------------------------------------------ packed Option<Outcome> (game.rs:9-66)
// 0 = None, else kind<<6 | turns, kind 1 = Lose, 2 = Draw, 3 = Win.
__device__ __forceinline__ uint32_t sol_kind(uint32_t s) { return s >> 6; }
__device__ __forceinline__ uint32_t sol_reversed(uint32_t s) { // game.rs:28-35; s != 0
    return ((4u - (s >> 6)) << 6) | (((s & 63u) + 1u) & 63u);
}
// Monotone key of the Ord impl (game.rs:46-60) extended to Option (None lowest): Win prefers FEWER
// turns, Draw and Lose prefer MORE.
__device__ __forceinline__ uint32_t sol_key(uint32_t s) {
    if (s == 0u) return 0u;
    return (s >> 6) == SYN_KIND_WIN ? ((3u << 6) | (63u - (s & 63u))) : s;
}
__device__ __forceinline__ uint32_t sol_from_key(uint32_t k) {
    if (k == 0u) return 0u;
    return (k >> 6) == SYN_KIND_WIN ? ((3u << 6) | (63u - (k & 63u))) : k;
}
__device__ __forceinline__ int sol_index(uint32_t s) { return (int)(s >> 6) - 1; } // Lose 0, Draw 1, Win 2 (mcts.rs:10-18)

} // namespace eng

namespace tpx {

enum { K_NONE = 0, K_LEAF = 1, K_TERMINAL = 2, K_INIT = 4 /* flag: the construction visit of MCTS::with_capacity */ };

struct Pend { // what descend leaves for finish
    uint32_t kind;  // K_*
    uint32_t id;    // K_LEAF: the expanded node; K_TERMINAL: the proven node
    uint32_t fc;    // K_LEAF: first child; K_TERMINAL: the node's packed solution
    uint32_t lc;    // K_LEAF: legal mask | csol2 << 9 (2 bits per column: 0 none / 1 Lose(0) / 2 Draw(0))
    uint32_t depth; // level of `id` (root = 0): path[0 .. depth-1] holds the ids of levels 1 .. depth
};

struct RoundCnt { uint32_t levels, scanned, expansions, created, bp_levels, leaf_evals, explores; };

struct RootOut {
    float pi[9], visits[9];
    uint32_t child_sol[9];
    float q0, q1, q2;
    uint32_t root_sol, legal;
    int best_action;
};

} // namespace tpx
