// c4.cuh — Connect4 9x7 on packed bitboards, device side.
//
// Replaces study-connect4/src/connect4.rs:15-83 (masks, `won`) and :173-258 (`Game` impl) of the
// reference.  Bit layout is the reference's: bit = row + 7*col, 63 bits used, no sentinel row; the
// masks alone stop a run from wrapping into the next column.  Only the two bitboards are kept:
// `height[col]` is popcount of the column's 7 bits (stones stack from the bottom) and the player
// to move is the parity of the stone count, so a position is 16 bytes in registers.
#pragma once
#include <stdint.h>
#include "devport.cuh"

namespace c4 {

constexpr int W = 9;
constexpr int H = 7;

__host__ __device__ constexpr uint64_t fab_row() {
    uint64_t r = 0;
    for (int c = 0; c < 9; ++c) r |= 1ull << (7 * c);
    return r;
}
constexpr uint64_t ROW0 = fab_row();                 // bottom cell of every column
constexpr uint64_t ALL = 0x7fffffffffffffffull;      // the 63 cells
constexpr uint64_t COLS05 = (1ull << 42) - 1;        // columns 0..5
constexpr uint64_t V_MASK = ROW0 * 0x0full;          // rows 0..3
constexpr uint64_t H_MASK = COLS05;
constexpr uint64_t D2_MASK = COLS05 & (ROW0 * 0x0full); // rows 0..3  (up-right diagonal starts)
constexpr uint64_t D1_MASK = COLS05 & (ROW0 * 0x78ull); // rows 3..6  (down-right diagonal starts)

// connect4.rs:77-83
// bb & bb>>s & bb>>2s & bb>>3s == t & t>>2s with t = bb & bb>>s: the same bits with two shifts per direction
// instead of three (this is the inner loop of every rollout ply).
__device__ __forceinline__ bool won(uint64_t bb) {
    const uint64_t t1 = bb & (bb >> 6), t2 = bb & (bb >> 8), t3 = bb & (bb >> 7), t4 = bb & (bb >> 1);
    uint64_t d1 = t1 & (t1 >> 12) & D1_MASK;
    uint64_t d2 = t2 & (t2 >> 16) & D2_MASK;
    uint64_t h = t3 & (t3 >> 14) & H_MASK;
    uint64_t v = t4 & (t4 >> 2) & V_MASK;
    return (d1 | d2 | h | v) != 0;
}

// All cells e such that `p | e` contains four in a row THROUGH e — the mover's winning cells for every
// column at once (Connect4::step + won, connect4.rs:77-83, 221-233, evaluated for the nine candidate
// moves of visit(), mcts.rs:384-397, in ~1/3 of the instructions of nine won() calls).  Stones are moved
// one neighbour step at a time; without a sentinel row each step masks the row that would wrap into the
// next column.  Equal to won(p | e) whenever p itself has no line yet (checked exhaustively against won()
// on random positions: tests/test_host_and_cabi.py mirrors the algebra, the GPU parity suite the kernel).
constexpr uint64_t ROW6 = ROW0 << 6;
__device__ __forceinline__ uint64_t st_up(uint64_t x) { return (x & ~ROW6) << 1; }
__device__ __forceinline__ uint64_t st_right(uint64_t x) { return (x << 7) & ALL; }
__device__ __forceinline__ uint64_t st_left(uint64_t x) { return x >> 7; }
__device__ __forceinline__ uint64_t st_ur(uint64_t x) { return ((x & ~ROW6) << 8) & ALL; }
__device__ __forceinline__ uint64_t st_dl(uint64_t x) { return (x & ~ROW0) >> 8; }
__device__ __forceinline__ uint64_t st_dr(uint64_t x) { return ((x & ~ROW0) << 6) & ALL; }
__device__ __forceinline__ uint64_t st_ul(uint64_t x) { return (x & ~ROW6) >> 6; }
#define SYN_LINE(F, B)                                                                  \
    {                                                                                   \
        uint64_t P1 = B(p), P2 = B(P1), P3 = B(P2), M1 = F(p), M2 = F(M1), M3 = F(M2);  \
        w |= (P1 & P2 & P3) | (M1 & P1 & P2) | (M2 & M1 & P1) | (M3 & M2 & M1);         \
    }
__device__ __forceinline__ uint64_t winning_cells(uint64_t p) {
    uint64_t b1 = st_up(p), b2 = st_up(b1), b3 = st_up(b2);
    uint64_t w = b1 & b2 & b3; // vertical: three of the mover's stones right below
    SYN_LINE(st_right, st_left)
    SYN_LINE(st_ur, st_dl)
    SYN_LINE(st_dr, st_ul)
    return w;
}
#undef SYN_LINE

// Bit of the lowest empty cell of column `col` (column must not be full): stones are bottom-
// contiguous, so adding the column's bottom bit carries up to the first hole.
__device__ __forceinline__ uint64_t drop_bit(uint64_t occ, int col) {
    uint64_t bottom = 1ull << (7 * col);
    return (occ + bottom) & (0x7full << (7 * col));
}

__device__ __forceinline__ int height(uint64_t occ, int col) { return __popcll((occ >> (7 * col)) & 0x7full); }

// 9-bit mask of columns with room (FreeColumns, connect4.rs:138-161), serial form.
__device__ __forceinline__ uint32_t legal_mask(uint64_t occ) {
    uint32_t m = 0;
#pragma unroll
    for (int c = 0; c < 9; ++c) m |= (uint32_t)(((~occ) >> (7 * c + 6)) & 1ull) << c;
    return m;
}

// Game::step (connect4.rs:221-233): the mover's stone goes onto my_bb, then the boards swap.
// Returns 0 = not over, SOL_LOSE0 (the mover won: the new player to move has lost) or SOL_DRAW0.
constexpr uint32_t SOL_LOSE0 = 1u << 6;
constexpr uint32_t SOL_DRAW0 = 2u << 6;
constexpr uint32_t SOL_WIN0 = 3u << 6;

__device__ __forceinline__ uint32_t step(uint64_t& my, uint64_t& op, int col) {
    uint64_t occ = my | op;
    uint64_t bit = drop_bit(occ, col);
    uint64_t mover = my | bit;
    my = op;
    op = mover;
    if (won(mover)) return SOL_LOSE0;
    if ((occ | bit) == ALL) return SOL_DRAW0;
    return 0u;
}

// Game::features (connect4.rs:237-258), flattened row*9+col: +1 mine, -1 theirs, -0.1 empty,
// +0.1 on the next playable cell of every column with room.
__device__ __forceinline__ float feature(uint64_t my, uint64_t op, int idx) {
    int row = idx / 9, col = idx - row * 9;
    uint64_t bit = 1ull << (row + 7 * col);
    uint64_t occ = my | op;
    if (my & bit) return 1.0f;
    if (op & bit) return -1.0f;
    return (height(occ, col) == row) ? 0.1f : -0.1f;
}

} // namespace c4
