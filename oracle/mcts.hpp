// oracle/mcts.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of synthesis/src/mcts.rs:29-489 (`MCTS`) and of the evaluator's frozen
// baseline synthesis/src/evaluator.rs:233-534 (`FrozenMCTS`).  Each function cites the lines
// it follows.  f32 arithmetic is written in the reference's evaluation order and this file
// must be compiled with -ffp-contract=off (Rust never contracts a*b+c into an FMA).
//
// Two deliberate, documented differences from a literal transcription:
//  * exp / ln go through syn_expf / syn_logf (include/syn_detmath.h) unless
//    TreeOptions::libm is set, because bit-exact GPU/CPU trees need one definition of both
//    (the reference uses the platform libm: mcts.rs:364,418; evaluator.rs:420,465).
//  * Fpu::Func (config.rs:25) cannot be represented; SYN_FPU_NORMAL draws Normal(a,b) from a
//    seeded stream instead of thread_rng (study-connect4/src/main.rs:43-47), and Dirichlet noise
//    draws from a seeded stream instead of thread_rng (mcts.rs:236).
// TreeOptions::legacy reproduces the semantics the reference's stale test constants were
// recorded under (solved children scored WITHOUT the exploration term, auto-extend ignored):
// nodes.len() == 311 / 69 / 1533 at mcts.rs:732,781,830.  HEAD semantics give 244 / 69 / 1467.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "../include/syn_detmath.h"
#include "../include/syn_sampling.h"
#include "../include/synthesis_b200.h"
#include "game.hpp"
#include "rng.hpp"

namespace orc {

struct TreeOptions {
    bool legacy = false; // pre-HEAD semantics (see header comment)
    bool libm = false;   // use std::exp / std::log instead of syn_expf / syn_logf
};

template <class G>
struct Policy {
    virtual ~Policy() {}
    // policies/traits.rs:4-6: (logits[N], outcome probabilities [Lose, Draw, Win])
    virtual void eval(const G& g, float logits[9], float probs[3]) = 0;
};

struct Counters {
    uint64_t explores = 0, leaf_evals = 0, trees = 0, nodes = 0, select_levels = 0, children_scanned = 0,
             expansions = 0, children_created = 0, backprop_levels = 0, rollout_plies = 0;
    void add(const Counters& o) {
        explores += o.explores; leaf_evals += o.leaf_evals; trees += o.trees; nodes += o.nodes;
        select_levels += o.select_levels; children_scanned += o.children_scanned; expansions += o.expansions;
        children_created += o.children_created; backprop_levels += o.backprop_levels; rollout_plies += o.rollout_plies;
    }
};

static inline float f_exp(float x, bool libm) { return libm ? std::exp(x) : syn_expf(x); }
static inline float f_ln(float x, bool libm) { return libm ? std::log(x) : syn_logf(x); }

// ------------------------------------------------------------------------- MCTS (mcts.rs)
template <class G>
struct MCTS {
    // mcts.rs:29-39
    struct Node {
        uint32_t parent = 0, first_child = 0;
        uint8_t num_children = 0;
        G game;
        Outcome solution; // kind == NONE <=> Option::None
        uint8_t action = 0;
        float action_prob = 0.0f;
        float outcome_probs[3] = {0.0f, 0.0f, 0.0f};
        float num_visits = 0.0f;
        float q() const { return (outcome_probs[2] - outcome_probs[0]) / num_visits; } // mcts.rs:42-44
        bool is_unvisited() const { return num_children == 0 && !solution.is_some(); } // mcts.rs:71-73
        uint32_t last_child() const { return first_child + num_children; }
    };

    std::vector<Node> nodes;
    uint32_t root = 0;
    syn_mcts_cfg cfg;
    TreeOptions opt;
    Policy<G>* policy;
    StdRng* noise_rng; // Dirichlet noise stream (may be null when noise_kind != DIRICHLET)
    StdRng* fpu_rng;   // Normal FPU stream (may be null unless fpu_kind == NORMAL)
    Counters* cnt;
    Counters local_cnt;

    // mcts.rs:123-137
    MCTS(size_t capacity, const syn_mcts_cfg& c, Policy<G>* p, const G& game, TreeOptions o = TreeOptions(),
         StdRng* noise = nullptr, StdRng* fpu = nullptr, Counters* counters = nullptr)
        : cfg(c), opt(o), policy(p), noise_rng(noise), fpu_rng(fpu), cnt(counters ? counters : &local_cnt) {
        nodes.reserve(capacity);
        Node r;
        r.game = game;
        nodes.push_back(r);
        cnt->trees += 1;
        float op[3];
        bool any_solved;
        uint32_t id = visit(root, op, any_solved);
        backprop(id, op, any_solved);
        add_root_noise();
    }

    // mcts.rs:139-147
    void explore_n(size_t n) {
        for (size_t i = 0; i < n; ++i) {
            if (nodes[root].solution.is_some()) break;
            explore();
        }
    }

    // mcts.rs:111-121
    static int exploit(size_t explores, const syn_mcts_cfg& c, Policy<G>* p, const G& game, uint32_t action_selection,
                       TreeOptions o = TreeOptions(), Counters* counters = nullptr) {
        MCTS m(explores + 1, c, p, game, o, nullptr, nullptr, counters);
        m.explore_n(explores);
        m.finish();
        return m.best_action(action_selection);
    }

    void finish() { cnt->nodes += nodes.size(); }

    // mcts.rs:174-211
    void target_policy(float pi[9]) const {
        for (int i = 0; i < 9; ++i) pi[i] = 0.0f;
        float total = 0.0f;
        const Node& r = nodes[root];
        if (r.num_visits == 1.0f) {
            if (r.solution.kind == WIN) {
                for (uint32_t c = r.first_child; c < r.last_child(); ++c) {
                    float v = nodes[c].solution.kind == LOSE ? 1.0f : 0.0f;
                    pi[nodes[c].action] = v;
                    total += v;
                }
            } else {
                for (uint32_t c = r.first_child; c < r.last_child(); ++c) {
                    pi[nodes[c].action] = 1.0f;
                    total += 1.0f;
                }
            }
        } else {
            for (uint32_t c = r.first_child; c < r.last_child(); ++c) {
                float v = nodes[c].num_visits;
                pi[nodes[c].action] = v;
                total += v;
            }
        }
        for (int i = 0; i < 9; ++i) pi[i] /= total;
    }

    // mcts.rs:213-225
    void target_q(float q[3]) const {
        const Node& r = nodes[root];
        if (r.solution.is_some()) {
            q[0] = q[1] = q[2] = 0.0f;
            q[r.solution.index()] = 1.0f;
        } else {
            for (int i = 0; i < 3; ++i) q[i] = r.outcome_probs[i] / r.num_visits;
        }
    }

    // mcts.rs:229-269
    void add_root_noise() {
        Node& r = nodes[root];
        if (cfg.noise_kind == SYN_NOISE_NONE) return;
        if (r.num_children < 2) return;
        float w = cfg.noise_weight;
        if (cfg.noise_kind == SYN_NOISE_EQUAL) {
            float noise = 1.0f / (float)r.num_children;
            for (uint32_t c = r.first_child; c < r.last_child(); ++c)
                nodes[c].action_prob = nodes[c].action_prob * (1.0f - w) + w * noise;
        } else {
            float noise[9];
            syn_dirichlet(*noise_rng, cfg.noise_alpha, (int)r.num_children, noise);
            for (uint32_t c = r.first_child, i = 0; c < r.last_child(); ++c, ++i)
                nodes[c].action_prob = nodes[c].action_prob * (1.0f - w) + w * noise[i];
        }
    }

    // mcts.rs:273-294.  Key = Option<(f32,f32)> compared with `>`; first child always wins over None.
    int best_action(uint32_t action_selection) const {
        const Node& r = nodes[root];
        int best = -1;
        bool have = false;
        float b0 = 0.0f, b1 = 0.0f;
        for (uint32_t c = r.first_child; c < r.last_child(); ++c) {
            const Node& ch = nodes[c];
            float v0, v1;
            if (ch.solution.kind == WIN) { v0 = 0.0f; v1 = (float)ch.solution.turns; }
            else if (ch.solution.kind == NONE) { v0 = 1.0f; v1 = action_selection == SYN_ACTION_Q ? -ch.q() : ch.num_visits; }
            else if (ch.solution.kind == DRAW) { v0 = 2.0f; v1 = -(float)ch.solution.turns; }
            else { v0 = 3.0f; v1 = -(float)ch.solution.turns; }
            // tuple partial_cmp: first elements decide unless equal; NaN second => not greater
            bool greater = !have || (v0 > b0) || (v0 == b0 && v1 > b1);
            if (greater) { have = true; b0 = v0; b1 = v1; best = ch.action; }
        }
        return best;
    }

    // mcts.rs:296-306
    Outcome solution(int action) const {
        const Node& r = nodes[root];
        for (uint32_t c = r.first_child; c < r.last_child(); ++c)
            if (nodes[c].action == (uint8_t)action) return nodes[c].solution;
        return Outcome::none();
    }

    // mcts.rs:310-325
    void explore() {
        cnt->explores += 1;
        uint32_t id = root;
        for (;;) {
            const Node& n = nodes[id];
            if (n.solution.is_some()) {
                float op[3] = {0.0f, 0.0f, 0.0f};
                op[n.solution.index()] = 1.0f;
                backprop(id, op, true);
                return;
            } else if (n.is_unvisited()) {
                float op[3];
                bool any_solved;
                uint32_t leaf = visit(id, op, any_solved);
                backprop(leaf, op, any_solved);
                return;
            } else {
                id = select_best_child(id);
            }
        }
    }

    // mcts.rs:327-341: Option<f32> compared with strict `>`; Some(x) > None always.
    uint32_t select_best_child(uint32_t pid) {
        const Node& p = nodes[pid];
        cnt->select_levels += 1;
        cnt->children_scanned += p.num_children;
        uint32_t best = 0;
        bool have = false;
        float bv = 0.0f;
        for (uint32_t c = p.first_child; c < p.last_child(); ++c) {
            const Node& ch = nodes[c];
            float value;
            if (opt.legacy && ch.solution.is_some()) {
                value = ch.solution.reversed().value();
            } else {
                float q = exploit_value(p, ch);
                float u = explore_value(p, ch);
                value = q + u;
            }
            if (!have || value > bv) { have = true; best = c; bv = value; }
        }
        return best;
    }

    // mcts.rs:343-359
    float exploit_value(const Node& p, const Node& ch) {
        if (ch.solution.is_some()) {
            return cfg.select_solved_nodes ? ch.solution.reversed().value() : -std::numeric_limits<float>::infinity();
        } else if (ch.num_children == 0) {
            if (cfg.fpu_kind == SYN_FPU_CONST) return cfg.fpu_a;
            if (cfg.fpu_kind == SYN_FPU_PARENT_Q) return p.q();
            return syn_normal(*fpu_rng, cfg.fpu_a, cfg.fpu_b);
        } else {
            return -ch.q();
        }
    }

    // mcts.rs:361-372
    float explore_value(const Node& p, const Node& ch) const {
        if (cfg.exploration_kind == SYN_EXPLORATION_UCT) {
            float visits = std::sqrt(cfg.c * f_ln(p.num_visits, opt.libm));
            return visits / std::sqrt(ch.num_visits);
        } else {
            float visits = std::sqrt(p.num_visits);
            return cfg.c * ch.action_prob * visits / (1.0f + ch.num_visits);
        }
    }

    // mcts.rs:374-427
    uint32_t visit(uint32_t id, float op[3], bool& any_solved_out) {
        uint32_t first_child = (uint32_t)nodes.size();
        if (nodes[id].solution.is_some()) {
            op[0] = op[1] = op[2] = 0.0f;
            op[nodes[id].solution.index()] = 1.0f;
            any_solved_out = true;
            return id;
        }
        G game = nodes[id].game;
        int acts[9];
        int na = game.actions(acts);
        bool any_solved = false;
        for (int i = 0; i < na; ++i) {
            Node ch;
            ch.game = game;
            bool over = ch.game.step(acts[i]);
            if (over) {
                any_solved = true;
                ch.solution = Outcome::from_f32(ch.game.reward(ch.game.player()));
            }
            ch.parent = id;
            ch.action = (uint8_t)acts[i];
            ch.action_prob = 1.0f;
            nodes.push_back(ch);
        }
        cnt->expansions += 1;
        cnt->children_created += na;
        nodes[id].first_child = first_child;
        nodes[id].num_children = (uint8_t)na;
        uint32_t last_child = first_child + na;

        if (cfg.auto_extend && !opt.legacy && na == 1) {
            return visit(first_child, op, any_solved_out);
        }
        float logits[9];
        cnt->leaf_evals += 1;
        policy->eval(game, logits, op);
        // stable softmax over the legal children (mcts.rs:409-423)
        float max_logit = -std::numeric_limits<float>::infinity();
        for (uint32_t c = first_child; c < last_child; ++c) {
            float logit = logits[nodes[c].action];
            max_logit = std::fmax(max_logit, logit); // f32::max ignores NaN like fmaxf
            nodes[c].action_prob = logit;
        }
        float total = 0.0f;
        for (uint32_t c = first_child; c < last_child; ++c) {
            nodes[c].action_prob = f_exp(nodes[c].action_prob - max_logit, opt.libm);
            total += nodes[c].action_prob;
        }
        for (uint32_t c = first_child; c < last_child; ++c) nodes[c].action_prob /= total;
        any_solved_out = any_solved;
        return id;
    }

    // mcts.rs:429-488
    void backprop(uint32_t leaf, float op_in[3], bool solved) {
        float op[3] = {op_in[0], op_in[1], op_in[2]};
        uint32_t id = leaf;
        for (;;) {
            cnt->backprop_levels += 1;
            uint32_t parent = nodes[id].parent;
            if (cfg.solve && solved) {
                bool all_solved = true;
                Outcome best = nodes[id].solution;
                for (uint32_t c = nodes[id].first_child; c < nodes[id].last_child(); ++c) {
                    Outcome s = nodes[c].solution.is_some() ? nodes[c].solution.reversed() : Outcome::none();
                    all_solved = all_solved && s.is_some();
                    best = outcome_max(best, s);
                }
                Node& n = nodes[id];
                bool correct = cfg.correct_values_on_solve != 0;
                if (best.kind == WIN) {
                    n.solution = best;
                    if (correct) {
                        for (int i = 0; i < 3; ++i) op[i] = -n.outcome_probs[i];
                        op[2] += n.num_visits + 1.0f;
                    }
                } else if (best.is_some() && all_solved) {
                    n.solution = best;
                    if (correct) {
                        for (int i = 0; i < 3; ++i) op[i] = -n.outcome_probs[i];
                        if (best.kind == DRAW) op[1] += n.num_visits + 1.0f;
                        else op[0] += n.num_visits + 1.0f;
                    }
                } else {
                    solved = false;
                }
            }
            Node& n = nodes[id];
            for (int i = 0; i < 3; ++i) n.outcome_probs[i] += op[i];
            n.num_visits += 1.0f;
            if (id == root) break;
            float t = op[0]; op[0] = op[2]; op[2] = t;
            id = parent;
        }
    }
};

// ------------------------------------------------------------------ FrozenMCTS (evaluator.rs:233-534)
template <class G>
struct FrozenMCTS {
    struct Node {
        uint32_t parent = 0, first_child = 0;
        uint8_t num_children = 0;
        G game;
        Outcome solution;
        uint8_t action = 0;
        float action_prob = 0.0f, cum_value = 0.0f, num_visits = 0.0f;
        bool is_unvisited() const { return num_children == 0 && !solution.is_some(); }
        bool is_visited() const { return num_children != 0; }
        uint32_t last_child() const { return first_child + num_children; }
    };
    std::vector<Node> nodes;
    uint32_t root = 0;
    syn_mcts_cfg cfg;
    TreeOptions opt;
    Policy<G>* policy;
    Counters* cnt;
    Counters local_cnt;
    bool unsupported = false; // the reference panics (evaluator.rs:410,424)

    // evaluator.rs:320-333
    FrozenMCTS(size_t capacity, const syn_mcts_cfg& c, Policy<G>* p, const G& game, TreeOptions o = TreeOptions(),
               Counters* counters = nullptr)
        : cfg(c), opt(o), policy(p), cnt(counters ? counters : &local_cnt) {
        nodes.reserve(capacity);
        Node r;
        r.game = game;
        nodes.push_back(r);
        cnt->trees += 1;
        bool any_solved;
        float v = visit(root, any_solved);
        backprop(root, v, any_solved);
    }
    // evaluator.rs:308-318
    static int exploit(size_t explores, const syn_mcts_cfg& c, Policy<G>* p, const G& game, uint32_t action_selection,
                       TreeOptions o = TreeOptions(), Counters* counters = nullptr) {
        FrozenMCTS m(explores + 1, c, p, game, o, counters);
        m.explore_n(explores);
        m.finish();
        return m.best_action(action_selection);
    }
    void finish() { cnt->nodes += nodes.size(); }
    // evaluator.rs:529-533 (no root-solved early exit)
    void explore_n(size_t n) {
        for (size_t i = 0; i < n; ++i) explore();
    }
    // evaluator.rs:356-380
    int best_action(uint32_t action_selection) const {
        const Node& r = nodes[root];
        int best = -1;
        float bv = -std::numeric_limits<float>::infinity();
        for (uint32_t c = r.first_child; c < r.last_child(); ++c) {
            const Node& ch = nodes[c];
            if (ch.is_unvisited()) continue;
            float value;
            if (ch.solution.kind == WIN) value = -std::numeric_limits<float>::infinity();
            else if (ch.solution.kind == DRAW) value = 1e6f;
            else if (ch.solution.kind == LOSE) value = std::numeric_limits<float>::infinity();
            else value = action_selection == SYN_ACTION_Q ? -ch.cum_value / ch.num_visits : ch.num_visits;
            if (best < 0 || value > bv) { bv = value; best = ch.action; }
        }
        return best;
    }
    // evaluator.rs:382-397
    void explore() {
        cnt->explores += 1;
        uint32_t id = root;
        for (;;) {
            const Node& n = nodes[id];
            if (n.solution.is_some()) {
                backprop(id, n.solution.value(), true);
                return;
            } else if (n.is_unvisited()) {
                bool any_solved;
                float v = visit(id, any_solved);
                backprop(id, v, any_solved);
                return;
            } else {
                id = select_best_child(id);
            }
        }
    }
    // evaluator.rs:399-438
    uint32_t select_best_child(uint32_t pid) {
        const Node& p = nodes[pid];
        cnt->select_levels += 1;
        cnt->children_scanned += p.num_children;
        bool have = false;
        uint32_t best = 0;
        float bv = -std::numeric_limits<float>::infinity();
        for (uint32_t c = p.first_child; c < p.last_child(); ++c) {
            const Node& ch = nodes[c];
            float value;
            if (ch.is_unvisited()) {
                if (cfg.fpu_kind != SYN_FPU_CONST) unsupported = true;
                value = cfg.fpu_a + ch.action_prob;
            } else {
                float q = ch.solution.is_some() ? ch.solution.reversed().value() : -ch.cum_value / ch.num_visits;
                if (cfg.exploration_kind != SYN_EXPLORATION_UCT) unsupported = true;
                float visits = std::sqrt(cfg.c * f_ln(p.num_visits, opt.libm));
                float u = visits / std::sqrt(ch.num_visits);
                value = q + u;
            }
            if (!have || value > bv) { have = true; best = c; bv = value; }
        }
        return best;
    }
    // evaluator.rs:440-480: the policy is evaluated BEFORE the children are pushed
    float visit(uint32_t id, bool& any_solved_out) {
        uint32_t first_child = (uint32_t)nodes.size();
        G game = nodes[id].game;
        float logits[9], dist[3];
        cnt->leaf_evals += 1;
        policy->eval(game, logits, dist);
        int acts[9];
        int na = game.actions(acts);
        bool any_solved = false;
        float max_logit = -std::numeric_limits<float>::infinity();
        for (int i = 0; i < na; ++i) {
            Node ch;
            ch.game = game;
            bool over = ch.game.step(acts[i]);
            if (over) {
                any_solved = true;
                ch.solution = Outcome::from_f32(ch.game.reward(ch.game.player()));
            }
            float logit = logits[acts[i]];
            max_logit = std::fmax(max_logit, logit);
            ch.parent = id;
            ch.action = (uint8_t)acts[i];
            ch.action_prob = logit;
            nodes.push_back(ch);
        }
        cnt->expansions += 1;
        cnt->children_created += na;
        nodes[id].first_child = first_child;
        nodes[id].num_children = (uint8_t)na;
        float total = 0.0f;
        for (uint32_t c = first_child; c < first_child + na; ++c) {
            nodes[c].action_prob = f_exp(nodes[c].action_prob - max_logit, opt.libm);
            total += nodes[c].action_prob;
        }
        for (uint32_t c = first_child; c < first_child + na; ++c) nodes[c].action_prob /= total;
        any_solved_out = any_solved;
        return dist[2] - dist[0];
    }
    // evaluator.rs:482-527
    void backprop(uint32_t leaf, float value, bool solved) {
        uint32_t id = leaf;
        for (;;) {
            cnt->backprop_levels += 1;
            uint32_t parent = nodes[id].parent;
            if (cfg.solve && solved && !nodes[id].solution.is_some()) {
                bool all_solved = true;
                Outcome worst = Outcome::none();
                for (uint32_t c = nodes[id].first_child; c < nodes[id].last_child(); ++c) {
                    const Node& ch = nodes[c];
                    if (ch.is_unvisited() || !ch.solution.is_some()) all_solved = false;
                    else if (!worst.is_some() || outcome_cmp(ch.solution, worst) < 0) worst = ch.solution;
                }
                Node& n = nodes[id];
                if (worst.kind == LOSE) {
                    n.solution = Outcome::win(0);
                    value = -n.cum_value + (n.num_visits + 1.0f);
                } else if (n.is_visited() && all_solved) {
                    Outcome best_for_me = worst.reversed();
                    n.solution = best_for_me;
                    if (best_for_me.kind == DRAW) value = -n.cum_value;
                    else value = -n.cum_value - (n.num_visits + 1.0f);
                } else {
                    solved = false;
                }
            }
            Node& n = nodes[id];
            n.cum_value += value;
            n.num_visits += 1.0f;
            value = -value;
            if (id == root) break;
            id = parent;
        }
    }
};

} // namespace orc
