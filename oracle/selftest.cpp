// oracle/selftest.cpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Pins the oracle against every known-answer test the reference holds for this path
//  before anything is compared with it:
//   game.rs:95-141        Outcome ordering (3 tests)
//   connect4.rs:300-498   Connect4 rules, rewards, the 63-ply draw, all four win directions (7 tests)
//   mcts.rs:691-831       MCTS solver: best actions + nodes.len() 311 / 69 / 1533 (legacy semantics;
//                         HEAD semantics as written give 244 / 69 / 1467 — see mcts.hpp)
//   mcts.rs:833-868       root noise invariants
// plus RFC 7539's ChaCha20 zero-key block for the ChaCha core and SURVEY.md §8(c)'s cross-check
// vectors (a second, independent restatement made at survey time).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "selfplay.hpp"

using namespace orc;

static int failures = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) {                                                           \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);          \
            ++failures;                                                          \
        }                                                                        \
    } while (0)

extern "C" int orc_ttt_kat(int which, uint32_t flags, uint32_t* nodes_len, int* best_action_q, uint8_t child_solution[9], uint8_t* root_solution);

static syn_mcts_cfg eval_policy_cfg() { // study-connect4/src/main.rs:58-66
    syn_mcts_cfg c;
    std::memset(&c, 0, sizeof(c));
    c.exploration_kind = SYN_EXPLORATION_POLYNOMIAL_UCT;
    c.c = 3.0f;
    c.solve = c.correct_values_on_solve = c.select_solved_nodes = c.auto_extend = 1;
    c.fpu_kind = SYN_FPU_CONST;
    c.fpu_a = 1.0f;
    c.noise_kind = SYN_NOISE_NONE;
    return c;
}

static void test_outcome() { // game.rs:95-141
    Outcome W = Outcome::win(0), D = Outcome::draw(0), L = Outcome::lose(0), N = Outcome::none();
    CHECK(outcome_cmp(W, D) > 0); CHECK(outcome_cmp(W, L) > 0); CHECK(outcome_cmp(D, L) > 0);
    CHECK(outcome_cmp(D, W) < 0); CHECK(outcome_cmp(L, W) < 0); CHECK(outcome_cmp(L, D) < 0);
    CHECK(outcome_cmp(W, W) == 0); CHECK(outcome_cmp(D, D) == 0); CHECK(outcome_cmp(L, L) == 0);
    CHECK(outcome_cmp(W, N) > 0); CHECK(outcome_cmp(D, N) > 0); CHECK(outcome_cmp(L, N) > 0);
    CHECK(outcome_cmp(N, N) == 0);
    CHECK(outcome_max(outcome_max(W, D), L) == W);
    CHECK(outcome_max(N, L) == L);
    // turn tiebreaks (game.rs:49,53,57)
    CHECK(outcome_cmp(Outcome::win(1), Outcome::win(3)) > 0);
    CHECK(outcome_cmp(Outcome::draw(1), Outcome::draw(3)) < 0);
    CHECK(outcome_cmp(Outcome::lose(1), Outcome::lose(3)) < 0);
    CHECK(Outcome::win(2).reversed() == Outcome::lose(3));
    CHECK(Outcome::from_f32(-1.0f) == Outcome::lose(0));
    CHECK(Outcome::from_f32(0.0f) == Outcome::draw(0));
}

static void test_connect4() {
    { // connect4.rs:300-315 test_first_wins
        Connect4 g;
        const int m[7] = {0, 1, 0, 1, 0, 1, 0};
        for (int i = 0; i < 6; ++i) CHECK(!g.step(m[i]));
        CHECK(g.step(m[6]));
        CHECK(g.is_over()); CHECK(g.winner() == 0); CHECK(g.reward(g.player()) == -1.0f);
        CHECK(g.player() == 1); CHECK(g.reward(1) == -1.0f); CHECK(g.reward(0) == 1.0f);
    }
    { // connect4.rs:317-334 test_second_wins
        Connect4 g;
        const int m[8] = {0, 1, 2, 1, 2, 1, 2, 1};
        for (int i = 0; i < 7; ++i) CHECK(!g.step(m[i]));
        CHECK(g.step(m[7]));
        CHECK(g.winner() == 1); CHECK(g.reward(g.player()) == -1.0f); CHECK(g.player() == 0);
        CHECK(g.reward(1) == 1.0f); CHECK(g.reward(0) == -1.0f);
    }
    { // connect4.rs:337-447 test_draw
        Connect4 g;
        auto has = [&](int c) {
            int a[9];
            int n = g.actions(a);
            for (int i = 0; i < n; ++i)
                if (a[i] == c) return true;
            return false;
        };
        for (int pair = 0; pair < 4; ++pair) {
            int a = 2 * pair, b = 2 * pair + 1;
            const int seq[12] = {a, b, a, b, b, a, b, a, a, b, a, b};
            for (int i = 0; i < 12; ++i) CHECK(!g.step(seq[i]));
            CHECK(has(a)); CHECK(!g.step(a)); CHECK(!has(a));
            CHECK(has(b)); CHECK(!g.step(b)); CHECK(!has(b));
        }
        for (int i = 0; i < 6; ++i) CHECK(!g.step(8));
        CHECK(has(8));
        CHECK(g.step(8));
        CHECK(g.is_over()); CHECK(g.winner() < 0); CHECK(g.reward(0) == 0.0f); CHECK(g.reward(1) == 0.0f);
    }
    // connect4.rs:450-498 the four win directions
    for (int row = 0; row < 7; ++row) {
        uint64_t bb = (1ull << row) | (1ull << (row + 7)) | (1ull << (row + 14)) | (1ull << (row + 21));
        for (int i = 0; i < 6; ++i) { CHECK(Connect4::won(bb)); bb <<= 7; }
    }
    for (int col = 0; col < 9; ++col) {
        uint64_t bb = 0xfull << (7 * col);
        for (int i = 0; i < 4; ++i) { CHECK(Connect4::won(bb)); bb <<= 1; }
    }
    for (int row = 3; row < 7; ++row) {
        uint64_t bb = (1ull << row) | (1ull << (row + 6)) | (1ull << (row + 12)) | (1ull << (row + 18));
        for (int i = 0; i < 6; ++i) { CHECK(Connect4::won(bb)); bb <<= 7; }
    }
    for (int col = 0; col < 6; ++col) {
        uint64_t bb = (1ull << (7 * col)) | (1ull << (7 * (col + 1) + 1)) | (1ull << (7 * (col + 2) + 2)) | (1ull << (7 * (col + 3) + 3));
        for (int i = 0; i < 4; ++i) { CHECK(Connect4::won(bb)); bb <<= 1; }
    }
    // three in a row / wrap-around across the column boundary must not count
    CHECK(!Connect4::won(0x7ull));
    CHECK(!Connect4::won((1ull << 5) | (1ull << 6) | (1ull << 7) | (1ull << 8)));
}

static void test_rng() {
    { // RFC 7539 §2.3.2-style sanity: ChaCha20 with an all-zero key/nonce, block 0
        uint32_t key[8] = {0, 0, 0, 0, 0, 0, 0, 0}, out[16];
        chacha_block(key, 0, 20, out);
        CHECK(out[0] == 0xade0b876u); // bytes 76 b8 e0 ad
        CHECK(out[1] == 0x903df1a0u); // bytes a0 f1 3d 90
    }
    StdRng r0 = StdRng::seed_from_u64(0);
    CHECK(r0.key[0] == 0xf973f2ecu); CHECK(r0.key[1] == 0x45cdb581u); CHECK(r0.key[2] == 0x7346f087u); CHECK(r0.key[3] == 0xad6cad06u);
    const uint32_t w0[6] = {0xcd2c6f7fu, 0xbb2a3fb2u, 0x8e27697bu, 0xc6017c94u, 0xcf310a16u, 0x069dc102u};
    for (int i = 0; i < 6; ++i) CHECK(r0.next_u32() == w0[i]);
    StdRng r1 = StdRng::seed_from_u64(1);
    const uint32_t w1[4] = {0xd3301861u, 0xf9681a64u, 0xcc0d694au, 0xb0f4d125u};
    for (int i = 0; i < 4; ++i) CHECK(r1.next_u32() == w1[i]);
    StdRng r7 = StdRng::seed_from_u64(7);
    const uint32_t w7[4] = {0x6aa8fbbeu, 0x07c2e0e9u, 0x247e5f86u, 0x4e9d34e8u};
    for (int i = 0; i < 4; ++i) CHECK(r7.next_u32() == w7[i]);
    StdRng rg = StdRng::seed_from_u64(0);
    const int g9[12] = {7, 6, 4, 6, 7, 0, 6, 5, 7, 2, 1, 6};
    for (int i = 0; i < 12; ++i) CHECK((int)rg.gen_range_u8(9) == g9[i]);
}

static void test_ttt_kats() {
    uint32_t nodes;
    int best;
    uint8_t cs[9], rs;
    // legacy semantics reproduce the reference's recorded constants
    orc_ttt_kat(0, 1, &nodes, &best, cs, &rs);
    std::printf("ttt win : legacy nodes=%u best=%d\n", nodes, best);
    CHECK(nodes == 311); CHECK(best == 6);
    for (int a = 0; a < 9; ++a) if (a != 6) CHECK(cs[a] == 0); // mcts.rs:721-729
    orc_ttt_kat(1, 1, &nodes, &best, cs, &rs);
    std::printf("ttt loss: legacy nodes=%u best=%d\n", nodes, best);
    CHECK(nodes == 69);
    orc_ttt_kat(2, 1, &nodes, &best, cs, &rs);
    std::printf("ttt draw: legacy nodes=%u best=%d\n", nodes, best);
    CHECK(nodes == 1533); CHECK(best == 1); CHECK(cs[0] == 0);
    // HEAD semantics (what the source text says today)
    orc_ttt_kat(0, 0, &nodes, &best, cs, &rs);
    std::printf("ttt win : head   nodes=%u best=%d root=%02x\n", nodes, best, rs);
    CHECK(nodes == 244); CHECK(best == 6);
    for (int a = 0; a < 9; ++a) if (a != 6) CHECK(cs[a] == 0);
    CHECK((rs >> 6) == WIN);
    orc_ttt_kat(1, 0, &nodes, &best, cs, &rs);
    std::printf("ttt loss: head   nodes=%u best=%d root=%02x\n", nodes, best, rs);
    CHECK(nodes == 69); CHECK((rs >> 6) == LOSE);
    orc_ttt_kat(2, 0, &nodes, &best, cs, &rs);
    std::printf("ttt draw: head   nodes=%u best=%d root=%02x\n", nodes, best, rs);
    CHECK(nodes == 1467); CHECK(best == 1); CHECK(cs[0] == 0); CHECK((rs >> 6) == DRAW);
}

static void test_noise() { // mcts.rs:833-868
    StdRng rng = StdRng::seed_from_u64(0);
    RolloutPolicy<TicTacToe> policy(&rng);
    syn_mcts_cfg cfg = eval_policy_cfg();
    cfg.c = 2.0f;
    cfg.fpu_a = std::numeric_limits<float>::infinity();
    cfg.select_solved_nodes = 0;
    cfg.auto_extend = 0;
    TicTacToe game;
    MCTS<TicTacToe> mcts(1601, cfg, &policy, game);
    float total = 0.0f;
    const auto& r = mcts.nodes[mcts.root];
    for (uint32_t c = r.first_child; c < r.last_child(); ++c) { CHECK(mcts.nodes[c].action_prob > 0.0f); total += mcts.nodes[c].action_prob; }
    CHECK(std::fabs(total - 1.0f) < 1e-6f);
    StdRng rng2 = StdRng::seed_from_u64(0);
    mcts.noise_rng = &rng2;
    mcts.cfg.noise_kind = SYN_NOISE_DIRICHLET;
    mcts.cfg.noise_alpha = 1.0f;
    mcts.cfg.noise_weight = 0.25f;
    mcts.add_root_noise();
    total = 0.0f;
    for (uint32_t c = r.first_child; c < r.last_child(); ++c) { CHECK(mcts.nodes[c].action_prob > 0.0f); total += mcts.nodes[c].action_prob; }
    CHECK(std::fabs(total - 1.0f) < 1e-6f);
}

static void search_c4(const std::vector<int>& moves, uint32_t E, uint64_t seed, float visits[9], uint32_t* nodes, float rootop[3],
                      Outcome* rootsol, Outcome childsol[9], uint64_t* words) {
    Connect4 g;
    for (int m : moves) g.step(m);
    StdRng rng = StdRng::seed_from_u64(seed);
    RolloutPolicy<Connect4> policy(&rng);
    MCTS<Connect4> m(E + 1, eval_policy_cfg(), &policy, g);
    m.explore_n(E);
    const auto& r = m.nodes[m.root];
    for (int i = 0; i < 9; ++i) { visits[i] = 0; childsol[i] = Outcome::none(); }
    for (uint32_t c = r.first_child; c < r.last_child(); ++c) { visits[m.nodes[c].action] = m.nodes[c].num_visits; childsol[m.nodes[c].action] = m.nodes[c].solution; }
    *nodes = (uint32_t)m.nodes.size();
    for (int i = 0; i < 3; ++i) rootop[i] = r.outcome_probs[i];
    *rootsol = r.solution;
    *words = rng.words_drawn;
}

static void test_survey_vectors() { // SURVEY.md §8(c) cross-check vectors (not reference-pinned)
    float v[9], op[3];
    uint32_t nodes;
    Outcome rs, cs[9];
    uint64_t words;
    auto eq9 = [&](const int* e) { for (int i = 0; i < 9; ++i) if (v[i] != (float)e[i]) return false; return true; };
    search_c4({}, 100, 0, v, &nodes, op, &rs, cs, &words);
    { const int e[9] = {16, 9, 2, 24, 2, 31, 3, 11, 2}; CHECK(eq9(e)); CHECK(nodes == 910); CHECK(op[0] == 36 && op[1] == 0 && op[2] == 65); }
    search_c4({}, 800, 0, v, &nodes, op, &rs, cs, &words);
    { const int e[9] = {47, 38, 87, 76, 41, 343, 24, 125, 19}; CHECK(eq9(e)); CHECK(nodes == 7210); CHECK(op[0] == 339 && op[1] == 0 && op[2] == 462); }
    search_c4({}, 800, 1, v, &nodes, op, &rs, cs, &words);
    { const int e[9] = {84, 14, 26, 215, 42, 364, 12, 26, 17}; CHECK(eq9(e)); CHECK(op[0] == 335 && op[1] == 0 && op[2] == 466); }
    search_c4({4, 4, 3, 3}, 800, 0, v, &nodes, op, &rs, cs, &words);
    { const int e[9] = {3, 10, 10, 1, 3, 9, 1, 3, 2}; CHECK(eq9(e)); CHECK(nodes == 388); CHECK(rs == Outcome::win(3)); CHECK(cs[2] == Outcome::lose(2)); }
    search_c4({4, 3, 4, 3, 4, 3}, 800, 0, v, &nodes, op, &rs, cs, &words);
    CHECK(nodes == 10); CHECK(rs == Outcome::win(1)); CHECK(cs[4] == Outcome::lose(0)); CHECK(words == 3);

    // full run_game vectors
    syn_rollout_cfg rc;
    std::memset(&rc, 0, sizeof(rc));
    rc.random_actions_until = 1; rc.sample_actions_until = 0; rc.stop_games_when_solved = 0;
    rc.value_target_kind = SYN_VALUE_Q; rc.action_selection = SYN_ACTION_NUM_VISITS;
    rc.mcts = eval_policy_cfg(); rc.leaf_eval_kind = SYN_LEAF_ROLLOUT;
    struct V { uint64_t g; uint32_t E; const char* moves; uint64_t sum_nodes; };
    const V vec[] = {{0, 50, "743641716112556277605", 8130}, {1, 50, "542042434480676008355502302", 10374},
                     {2, 50, "15462382653550447302", 8021}, {0, 200, "7265355355022074473446406", 35632},
                     {1, 200, "547677656166233757781724300", 41111}};
    for (const V& x : vec) {
        rc.num_explores = x.E;
        LeafSource leaf;
        std::vector<GameTrace> tr;
        GatherResult r = gather_streams(rc, leaf, 0, x.g, 1, 1, TreeOptions(), &tr);
        std::string s;
        uint64_t sum = 0;
        for (size_t k = 0; k < tr[0].actions.size(); ++k) { s += (char)('0' + tr[0].actions[k]); sum += tr[0].tree_nodes[k]; }
        std::printf("run_game g=%llu E=%u -> %s  sum_nodes=%llu\n", (unsigned long long)x.g, x.E, s.c_str(), (unsigned long long)sum);
        CHECK(s == x.moves); CHECK(sum == x.sum_nodes);
        if (x.g == 0 && x.E == 50) {
            CHECK(tr[0].tree_nodes[0] == 460 && tr[0].tree_nodes[3] == 460);
            CHECK(r.buffer.vs[0][0] == 15.0f / 51.0f && r.buffer.vs[0][1] == 0.0f && r.buffer.vs[0][2] == 36.0f / 51.0f);
            const float e[9] = {.02f, .02f, .06f, .32f, .46f, .06f, .02f, .02f, .02f};
            for (int i = 0; i < 9; ++i) CHECK(std::fabs(r.buffer.pis[1][i] - e[i]) < 1e-6f);
        }
    }
}

int main() {
    test_outcome();
    test_connect4();
    test_rng();
    test_ttt_kats();
    test_noise();
    test_survey_vectors();
    if (failures) { std::printf("%d FAILURES\n", failures); return 1; }
    std::printf("oracle selftest: all checks passed\n");
    return 0;
}
