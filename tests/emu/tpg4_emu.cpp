// tests/emu/tpg4_emu.cpp — TEST INFRASTRUCTURE.  Steps the product's own device tree code
// (synthesis_b200/csrc/tpg4_tree.cuh, compiled as plain C++ through devport.cuh's SYN_HOST_EMU shim) on the CPU,
// one explore at a time exactly as selfplay_nn_tpg4_kernel's round does, and compares the finished tree NODE FOR NODE
// with the oracle's restatement of synthesis/src/mcts.rs (oracle/mcts.hpp): visit counts, outcome sums, priors,
// solutions, links — every bit.  Both sides evaluate leaves with the same deterministic policy object
// (RolloutPolicy on equal streams, or a hash of the position standing in for a network), so equal trees are required.
//
//   g++ -O2 -std=c++17 -ffp-contract=off -DSYN_HOST_EMU tpg4_emu.cpp -o tpg4_emu && ./tpg4_emu [positions] [explores]
//
// This checks the family-block layout (column slots, u16 visit counts, memoised -q, one-line blocks for only children), the
// solved walk, the chunked walk up and the owner-link fallback beyond PATH_CAP without a GPU; the -m gpu suite checks the
// same functions inside the kernels.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../oracle/mcts.hpp"
#include "../../oracle/policies.hpp"
#include "../../synthesis_b200/csrc/tpg4_tree.cuh"

using orc::Connect4;

// A stand-in for Connect4Net: logits and outcome probabilities are a hash of the position (any f32 values do: the
// tree code must treat them exactly as the oracle does).
struct HashPolicy : orc::Policy<Connect4> {
    uint64_t salt;
    explicit HashPolicy(uint64_t s) : salt(s) {}
    static uint64_t mix(uint64_t x) {
        x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
        return x;
    }
    void eval(const Connect4& g, float logits[9], float probs[3]) override {
        uint64_t h = mix(g.my_bb ^ mix(g.op_bb + salt));
        for (int i = 0; i < 9; ++i) {
            h = mix(h + i);
            logits[i] = ((float)(int)(h & 0xffff) - 32768.0f) / 8192.0f; // [-4, 4)
        }
        float e[3], t = 0.0f;
        for (int i = 0; i < 3; ++i) {
            h = mix(h + 77);
            e[i] = syn_expf(((float)(int)(h & 0xfff) - 2048.0f) / 1024.0f);
            t += e[i];
        }
        for (int i = 0; i < 3; ++i) probs[i] = e[i] / t;
    }
};

struct EmuTree {
    std::vector<uint4> nodes; // 8 x uint4 per 128-byte line
    uint32_t nn = 1, nb = 1;
    uint64_t root_occ = 0;
};

template <int FPU, int PATH_CAP, int CH>
static int emu_search(const syn_mcts_cfg& cfg, uint32_t explores, uint64_t my0, uint64_t op0, orc::Policy<Connect4>* pol, orc::StdRng* fpu_rng,
                      EmuTree& t) {
    const uint32_t blocks = 6u * (explores + 1u) + 64u; // lines
    t.nodes.assign(8 * (size_t)blocks, make_uint4(0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu)); // arenas are NOT zeroed between trees
    t.nn = 1; t.nb = 1; t.root_occ = my0 | op0;
    uint4* arena = t.nodes.data();
    uint32_t path[64];
    tp4::store_empty_root(arena);
    const float stop_vis = (float)(explores + 1u);
    auto draw = [&]() { return syn_normal(*fpu_rng, cfg.fpu_a, cfg.fpu_b); };
    for (;;) {
        tp4::Cur root = tp4::load_root(arena);
        if (root.vis >= stop_vis || tp4::link_sol(root.link) != 0u) break;
        tpx::Pend pd = {tpx::K_NONE, 0u, 0u, 0u, 0u};
        tpx::RoundCnt rc = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
        uint64_t my = my0, op = op0;
        int err = tp4::descend<FPU, 1, PATH_CAP>(cfg, blocks, draw, arena, t.nb, t.nn, root, my, op, pd, rc, path);
        if (err) return err;
        float v0, v1, v2;
        bool solved;
        if (pd.kind & tpx::K_LEAF) {
            float lg[9], pr[3];
            Connect4 g = Connect4::from_bitboards(my, op);
            pol->eval(g, lg, pr);
            tp4::write_children(arena, pd, lg);
            v0 = pr[0]; v1 = pr[1]; v2 = pr[2];
            solved = (pd.lc >> 9) != 0u;
        } else {
            int idx = eng::sol_index(pd.fc);
            v0 = idx == 0 ? 1.0f : 0.0f; v1 = idx == 1 ? 1.0f : 0.0f; v2 = idx == 2 ? 1.0f : 0.0f;
            solved = true;
        }
        tp4::prefetch_path<1, PATH_CAP>(arena, path, pd); // (a no-op on the CPU; the kernels call it right after the descent)
        tp4::backprop<1, PATH_CAP, CH>(cfg, arena, path, pd.depth, pd.id, v0, v1, v2, solved);
    }
    return 0;
}

static uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// node-for-node comparison; returns the number of differing fields.  Oracle node i <-> our id: the root is 0 <-> 0 and
// child j of a node whose children block is B is (B << 4) | j; blocks are handed out in expansion order like the Vec's pushes.
static int compare(const orc::MCTS<Connect4>& m, const EmuTree& t, const char* what, bool verbose) {
    int bad = 0;
    if (m.nodes.size() != t.nn) { if (verbose) std::printf("  %s: nodes.len() %zu vs %u\n", what, m.nodes.size(), t.nn); return 1; }
    const uint4* arena = t.nodes.data();
    std::vector<uint32_t> ours(m.nodes.size(), tp4::ROOT_ID);
    for (uint32_t i = 0; i < m.nodes.size(); ++i) {
        const auto& n = m.nodes[i];
        const uint32_t id = ours[i], k = tp4::id_slot(id);
        const uint32_t* l0 = tp4::line0(arena, tp4::id_line(id));
        const uint32_t link = l0[tp4::link_word(id)];
        const uint32_t vis = (id & tp4::SINGLE) ? (l0[0] & 0xffffu) : tp4::vis_of(l0[tp4::W_VIS + (k >> 1)], k);
        const uint32_t fc = tp4::link_fc(link), nch = tp4::link_nch(link), sol = tp4::link_sol(link);
        const uint32_t* st = l0 + tp4::stat_word(id);
        int bd = 0;
        bd += n.num_visits != (float)vis;                 // the select line's u16 copy
        bd += fbits(n.num_visits) != st[0];               // the stat record
        bd += fbits(n.outcome_probs[0]) != st[1]; bd += fbits(n.outcome_probs[1]) != st[2]; bd += fbits(n.outcome_probs[2]) != st[3];
        if (i != 0) { bd += fbits(n.action_prob) != l0[(id & tp4::SINGLE) ? 2u : tp4::W_PRIOR + k]; bd += ours[n.parent] != l0[tp4::owner_word(id)]; }
        bd += n.num_children != nch; bd += (n.num_children != 0) != (fc != 0); bd += n.solution.packed() != sol;
        if (n.num_visits > 0.0f) bd += fbits(-n.q()) != l0[tp4::nq_word(id)]; // the memo holds -q() of every node that has been backed up
        for (uint32_t j = 0; j < n.num_children; ++j)
            ours[n.first_child + j] = n.num_children == 1 ? (tp4::SINGLE | (fc << 4)) : ((fc << 4) | m.nodes[n.first_child + j].action);
        if (bd && verbose && bad < 5)
            std::printf("  %s: node %u (id %x) differs (%d fields): visits %g/%u sums %g %g %g / %g %g %g nch %u/%u sol %u/%u\n", what, i, id, bd, n.num_visits,
                        vis, n.outcome_probs[0], n.outcome_probs[1], n.outcome_probs[2], __uint_as_float(st[1]), __uint_as_float(st[2]), __uint_as_float(st[3]),
                        (unsigned)n.num_children, nch, (unsigned)n.solution.packed(), sol);
        bad += bd;
    }
    // the readout the driver uses (mcts.rs:174-225, 273-306)
    for (uint32_t sel_kind = 0; sel_kind < 2; ++sel_kind) {
        tpx::RootOut r;
        tp4::read_root(arena, t.root_occ, sel_kind, r);
        float pi[9], q[3];
        m.target_policy(pi);
        m.target_q(q);
        for (int k = 0; k < 9; ++k) bad += fbits(pi[k]) != fbits(r.pi[k]);
        bad += fbits(q[0]) != fbits(r.q0); bad += fbits(q[1]) != fbits(r.q1); bad += fbits(q[2]) != fbits(r.q2);
        bad += m.best_action(sel_kind) != r.best_action;
        bad += m.nodes[0].solution.packed() != r.root_sol;
        for (uint32_t c = m.nodes[0].first_child; c < m.nodes[0].last_child(); ++c) {
            bad += m.nodes[c].num_visits != r.visits[m.nodes[c].action];
            bad += m.nodes[c].solution.packed() != r.child_sol[m.nodes[c].action];
        }
    }
    return bad;
}

static double g_max_lines_per_explore = 0.0;

template <int FPU, int PATH_CAP, int CH>
static int run_case(const char* name, syn_mcts_cfg cfg, bool rollout, uint32_t explores, int positions, uint64_t seed0, uint32_t min_plies = 0,
                    uint32_t span_plies = 41) {
    int bad = 0, trees = 0;
    uint64_t total_nodes = 0;
    orc::StdRng walk = orc::StdRng::seed_from_u64(seed0 * 7919 + 13);
    for (int i = 0; i < positions; ++i) {
        // a random non-terminal position 0..40 plies into a game
        Connect4 g;
        uint32_t plies = min_plies + walk.gen_range_u8(span_plies);
        for (uint32_t k = 0; k < plies; ++k) {
            int acts[9];
            int n = g.actions(acts);
            Connect4 h = g;
            if (h.step(acts[walk.gen_range_u8((uint32_t)n)])) break;
            g = h;
        }
        const uint64_t seed = seed0 * 1000 + i;
        orc::StdRng r1 = orc::StdRng::seed_from_u64(seed), r2 = orc::StdRng::seed_from_u64(seed);
        orc::StdRng f1 = orc::StdRng::seed_from_u64(seed ^ 0x5555), f2 = orc::StdRng::seed_from_u64(seed ^ 0x5555);
        orc::RolloutPolicy<Connect4> ro1(&r1), ro2(&r2);
        HashPolicy h1(seed), h2(seed);
        orc::Policy<Connect4>* p1 = rollout ? (orc::Policy<Connect4>*)&ro1 : (orc::Policy<Connect4>*)&h1;
        orc::Policy<Connect4>* p2 = rollout ? (orc::Policy<Connect4>*)&ro2 : (orc::Policy<Connect4>*)&h2;
        syn_mcts_cfg c = cfg;
        c.noise_kind = SYN_NOISE_NONE; // root noise is not part of tpg4_tree.cuh
        orc::MCTS<Connect4> m(explores + 1, c, p1, g, orc::TreeOptions(), nullptr, &f1);
        m.explore_n(explores);
        EmuTree t;
        int err = emu_search<FPU, PATH_CAP, CH>(c, explores, g.my_bb, g.op_bb, p2, &f2, t);
        if (err) { std::printf("  %s: position %d: device error %d\n", name, i, err); ++bad; continue; }
        int b = compare(m, t, name, true);
        if (b) std::printf("  %s: position %d (ply %u): %d differing fields\n", name, i, plies, b);
        bad += b;
        ++trees;
        total_nodes += t.nn;
        const double lpe = (double)t.nb / (double)(m.cnt->explores + 1);
        if (lpe > g_max_lines_per_explore) g_max_lines_per_explore = lpe;
    }
    std::printf("%-58s %s  (%d trees, %llu nodes compared)\n", name, bad ? "FAIL" : "ok", trees, (unsigned long long)total_nodes);
    return bad;
}

int main(int argc, char** argv) {
    const int positions = argc > 1 ? std::atoi(argv[1]) : 40;
    const uint32_t explores = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 400;
    syn_mcts_cfg base;
    std::memset(&base, 0, sizeof(base));
    base.exploration_kind = SYN_EXPLORATION_POLYNOMIAL_UCT; base.c = 3.0f;
    base.solve = 1; base.correct_values_on_solve = 1; base.select_solved_nodes = 1; base.auto_extend = 1;
    base.fpu_kind = SYN_FPU_CONST; base.fpu_a = 1.0f; base.fpu_b = 0.0f;
    int bad = 0;
    // study-connect4/src/main.rs:58-66, rollout leaves and network-like leaves
    bad += run_case<SYN_FPU_CONST, 10, 2>("puct c=3 fpu=1 solver auto-extend, rollout leaves", base, true, explores, positions, 1);
    bad += run_case<SYN_FPU_CONST, 10, 2>("puct c=3 fpu=1 solver auto-extend, hashed-net leaves", base, false, explores, positions, 2);
    // the parent-link fallback of backprop beyond the path table, and every chunk size
    bad += run_case<SYN_FPU_CONST, 2, 1>("  ... PATH_CAP=2, chunk 1", base, false, explores, positions, 3);
    bad += run_case<SYN_FPU_CONST, 3, 4>("  ... PATH_CAP=3, chunk 4", base, true, explores, positions, 4);
    bad += run_case<SYN_FPU_CONST, 12, 3>("  ... PATH_CAP=12, chunk 3", base, false, explores, positions, 5);
    { syn_mcts_cfg c = base; c.fpu_kind = SYN_FPU_PARENT_Q;
      bad += run_case<SYN_FPU_PARENT_Q, 10, 2>("Fpu::ParentQ (memo of the parent)", c, false, explores, positions, 6);
      bad += run_case<SYN_FPU_PARENT_Q, 10, 2>("Fpu::ParentQ, rollout leaves", c, true, explores, positions, 7); }
    { syn_mcts_cfg c = base; c.fpu_kind = SYN_FPU_NORMAL; c.fpu_a = 1.0f; c.fpu_b = 0.1f;
      bad += run_case<SYN_FPU_NORMAL, 10, 2>("Fpu::Normal(1.0, 0.1) on a seeded stream (main.rs:43-47)", c, false, explores, positions, 8); }
    { syn_mcts_cfg c = base; c.exploration_kind = SYN_EXPLORATION_UCT; c.c = 2.0f; c.fpu_a = __uint_as_float(0x7f800000u); c.auto_extend = 0;
      bad += run_case<SYN_FPU_CONST, 10, 2>("uct c=2 fpu=inf, no auto-extend (main.rs:74-82)", c, true, explores, positions, 9); }
    { syn_mcts_cfg c = base; c.select_solved_nodes = 0;
      bad += run_case<SYN_FPU_CONST, 10, 2>("select_solved_nodes = false", c, true, explores, positions, 10); }
    { syn_mcts_cfg c = base; c.correct_values_on_solve = 0;
      bad += run_case<SYN_FPU_CONST, 10, 2>("correct_values_on_solve = false", c, false, explores, positions, 11); }
    { syn_mcts_cfg c = base; c.solve = 0;
      bad += run_case<SYN_FPU_CONST, 10, 2>("solve = false", c, true, explores, positions, 12); }
    { syn_mcts_cfg c = base; c.auto_extend = 0; c.fpu_a = 0.0f;
      bad += run_case<SYN_FPU_CONST, 10, 2>("no auto-extend, fpu = 0", c, false, explores, positions, 13); }
    // the ends of games: few columns with room, single-child chains (auto-extend), one-line families, proven roots
    bad += run_case<SYN_FPU_CONST, 10, 2>("late positions (40..62 plies), rollout leaves", base, true, explores, positions * 3, 14, 40, 23);
    bad += run_case<SYN_FPU_CONST, 10, 2>("late positions (40..62 plies), hashed-net leaves", base, false, explores, positions * 3, 15, 40, 23);
    std::printf("most arena lines used per executed explore (+1): %.3f (the engine provides 6 per explore + 64)\n", g_max_lines_per_explore);
    std::printf(bad ? "FAILED: %d differing fields\n" : "all trees identical to the oracle's, bit for bit\n", bad);
    return bad ? 1 : 0;
}
