// oracle/game.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of
//   synthesis/src/game.rs:9-66          Outcome, From<f32>, reversed, value, Ord
//   synthesis/src/mcts.rs:10-26         Outcome -> one-hot index / [f32;3]
//   study-connect4/src/connect4.rs:15-83,108-258   9x7 Connect4 on two bitboards + heights
//   synthesis/src/mcts.rs:499-686       the TicTacToe used by the reference's MCTS tests
// Pinned by: connect4.rs:300-498 (7 tests), game.rs:95-141 (3 tests) — restated in selftest.cpp.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>

namespace orc {

// ---------------------------------------------------------------- Outcome (game.rs:9-66)
enum Kind : uint8_t { NONE = 0, LOSE = 1, DRAW = 2, WIN = 3 };
struct Outcome {
    uint8_t kind = NONE; // NONE encodes Option::None
    uint32_t turns = 0;
    bool is_some() const { return kind != NONE; }
    static Outcome none() { return Outcome{}; }
    static Outcome win(uint32_t t) { return Outcome{WIN, t}; }
    static Outcome lose(uint32_t t) { return Outcome{LOSE, t}; }
    static Outcome draw(uint32_t t) { return Outcome{DRAW, t}; }
    // game.rs:16-26
    static Outcome from_f32(float v) { return v > 0.0f ? win(0) : (v < 0.0f ? lose(0) : draw(0)); }
    // game.rs:28-35
    Outcome reversed() const {
        if (kind == WIN) return lose(turns + 1);
        if (kind == LOSE) return win(turns + 1);
        return draw(turns + 1);
    }
    // game.rs:37-43
    float value() const { return kind == WIN ? 1.0f : (kind == DRAW ? 0.0f : -1.0f); }
    // mcts.rs:10-18: Lose=0, Draw=1, Win=2
    int index() const { return kind == LOSE ? 0 : (kind == DRAW ? 1 : 2); }
    bool operator==(const Outcome& o) const { return kind == o.kind && (kind == NONE || turns == o.turns); }
    uint8_t packed() const { return kind == NONE ? 0 : (uint8_t)((kind << 6) | (turns & 63u)); }
};

// game.rs:46-60 extended to Option<Outcome> (None < Some(_)).  Returns -1, 0, +1.
static inline int outcome_cmp(const Outcome& a, const Outcome& b) {
    if (!a.is_some() || !b.is_some()) return (int)a.is_some() - (int)b.is_some();
    auto c3 = [](uint32_t x, uint32_t y) { return x < y ? -1 : (x > y ? 1 : 0); };
    if (a.kind == WIN && b.kind == WIN) return c3(b.turns, a.turns); // fewer turns is greater
    if (a.kind == b.kind) return c3(a.turns, b.turns);               // Draw/Lose: more turns is greater
    return a.kind > b.kind ? 1 : -1;                                 // Win > Draw > Lose
}
// Ord::max(self, other): returns `other` unless self is strictly greater.
static inline Outcome outcome_max(const Outcome& a, const Outcome& b) { return outcome_cmp(a, b) > 0 ? a : b; }

// ---------------------------------------------------------------- Connect4 (connect4.rs)
struct Connect4 {
    static constexpr int N = 9;
    static constexpr int MAX_TURNS = 63;
    static constexpr int W = 9, H = 7;
    uint64_t my_bb = 0, op_bb = 0;
    uint8_t height[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint8_t player_ = 0; // 0 = Red (first), 1 = Black

    static constexpr uint64_t FAB_COL = 0x7full;
    static constexpr uint64_t fab_row() {
        uint64_t r = 0;
        for (int c = 0; c < 9; ++c) r |= 1ull << (7 * c);
        return r;
    }
    static constexpr uint64_t cols(int lo, int hi) {
        uint64_t r = 0;
        for (int c = lo; c <= hi; ++c) r |= FAB_COL << (7 * c);
        return r;
    }
    static constexpr uint64_t rows(int lo, int hi) {
        uint64_t r = 0;
        for (int k = lo; k <= hi; ++k) r |= fab_row() << k;
        return r;
    }
    // connect4.rs:70-83
    static bool won(uint64_t bb) {
        constexpr uint64_t D1 = cols(0, 5) & rows(3, 6), D2 = cols(0, 5) & rows(0, 3), HM = cols(0, 5), VM = rows(0, 3);
        uint64_t d1 = bb & (bb >> 6) & (bb >> 12) & (bb >> 18) & D1;
        uint64_t d2 = bb & (bb >> 8) & (bb >> 16) & (bb >> 24) & D2;
        uint64_t h = bb & (bb >> 7) & (bb >> 14) & (bb >> 21) & HM;
        uint64_t v = bb & (bb >> 1) & (bb >> 2) & (bb >> 3) & VM;
        return v + h + d1 + d2 > 0; // wrapping add like release-mode Rust; cannot overflow to 0 here
    }
    int player() const { return player_; }
    // connect4.rs:164-170: the only possible winner is the previous mover
    int winner() const { return won(op_bb) ? (player_ ^ 1) : -1; }
    bool is_over() const {
        if (winner() >= 0) return true;
        for (int c = 0; c < W; ++c)
            if (height[c] != H) return false;
        return true;
    }
    float reward(int pid) const {
        int w = winner();
        if (w < 0) return 0.0f;
        return w == pid ? 1.0f : -1.0f;
    }
    // connect4.rs:138-161,214-219: columns ascending with room
    int actions(int out[9]) const {
        int n = 0;
        for (int c = 0; c < W; ++c)
            if (height[c] < H) out[n++] = c;
        return n;
    }
    // connect4.rs:221-233
    bool step(int col) {
        my_bb ^= 1ull << (height[col] + 7 * col);
        height[col] += 1;
        uint64_t t = my_bb; my_bb = op_bb; op_bb = t;
        player_ ^= 1;
        return is_over();
    }
    // connect4.rs:237-258, flattened row*9+col (DIMS [1,1,7,9])
    void features(float s[63]) const {
        for (int row = 0; row < H; ++row)
            for (int col = 0; col < W; ++col) {
                uint64_t index = 1ull << (row + H * col);
                s[row * 9 + col] = (my_bb & index) ? 1.0f : ((op_bb & index) ? -1.0f : -0.1f);
            }
        for (int col = 0; col < W; ++col)
            if (height[col] < H) s[height[col] * 9 + col] = 0.1f;
    }
    // connect4.rs:116-121: Hash covers the two bitboards only
    uint64_t key_lo() const { return my_bb; }
    uint64_t key_hi() const { return op_bb; }
    bool operator==(const Connect4& o) const {
        return my_bb == o.my_bb && op_bb == o.op_bb && player_ == o.player_ && std::memcmp(height, o.height, 9) == 0;
    }
    // Rebuild a Connect4 from its bitboards (heights/player are functions of them for any
    // position reachable from new()).
    static Connect4 from_bitboards(uint64_t my, uint64_t op) {
        Connect4 g;
        g.my_bb = my; g.op_bb = op;
        int stones = 0;
        for (int c = 0; c < W; ++c) {
            int h = __builtin_popcountll(((my | op) >> (7 * c)) & FAB_COL);
            g.height[c] = (uint8_t)h;
            stones += h;
        }
        g.player_ = (uint8_t)(stones & 1);
        return g;
    }
};

// ---------------------------------------------------------------- TicTacToe (mcts.rs:499-686)
struct TicTacToe {
    static constexpr int N = 9;
    static constexpr int MAX_TURNS = 9;
    int8_t board[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1}; // -1 empty, 0 = X, 1 = O
    uint8_t player_ = 0;
    int turn = 0;
    int player() const { return player_; }
    bool won(int p) const {
        static const int L[8][3] = {{0, 1, 2}, {3, 4, 5}, {6, 7, 8}, {0, 3, 6}, {1, 4, 7}, {2, 5, 8}, {0, 4, 8}, {2, 4, 6}};
        for (auto& l : L)
            if (board[l[0]] == p && board[l[1]] == p && board[l[2]] == p) return true;
        return false;
    }
    bool is_over() const { return won(player_) || won(player_ ^ 1) || turn == 9; }
    float reward(int pid) const { return won(pid) ? 1.0f : (won(pid ^ 1) ? -1.0f : 0.0f); }
    int actions(int out[9]) const {
        int n = 0;
        for (int i = 0; i < 9; ++i)
            if (board[i] < 0) out[n++] = i;
        return n;
    }
    bool step(int a) {
        board[a] = (int8_t)player_;
        player_ ^= 1;
        turn += 1;
        return is_over();
    }
    uint64_t key_lo() const {
        uint64_t k = 0;
        for (int i = 0; i < 9; ++i) k = k * 3 + (uint64_t)(board[i] + 1);
        return k;
    }
    uint64_t key_hi() const { return player_; }
};

} // namespace orc
