// oracle/policies.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the reference's Policy implementations:
//   synthesis/src/policies/rollout.rs:8-31   RolloutPolicy
//   synthesis/src/policies/cache.rs:19-32    PolicyWithCache
//   study-connect4/src/policies.rs:13-59     Connect4Net (MLP 63-128-96-64-48-12, ReLU)
// The MLP arithmetic itself lives in libtorch (tch = "0.4.1", not vendored); here it is the
// plain fp32 definition in the loop order of slimnn/src/linear.rs:17-25 (bias first, then for
// each input, for each output: out += x * w, unfused).  No reference test fixes a network
// output ("parity unpinned"); the contract is BASELINE.json's 1e-3 abs/rel.
#pragma once
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "mcts.hpp"

namespace orc {

// policies/rollout.rs:8-31
template <class G>
struct RolloutPolicy : Policy<G> {
    StdRng* rng;
    Counters* cnt;
    RolloutPolicy(StdRng* r, Counters* c = nullptr) : rng(r), cnt(c) {}
    void eval(const G& game, float logits[9], float probs[3]) override {
        int player = game.player();
        G g = game;
        bool over = game.is_over();
        while (!over) {
            int acts[9];
            int n = g.actions(acts);
            uint32_t i = rng->gen_range_u8((uint32_t)n);
            over = g.step(acts[i]);
            if (cnt) cnt->rollout_plies += 1;
        }
        float r = g.reward(player);
        for (int i = 0; i < 9; ++i) logits[i] = 0.0f;
        probs[0] = probs[1] = probs[2] = 0.0f;
        if (r == 0.0f) probs[1] = 1.0f;
        else if (r < 0.0f) probs[0] = 1.0f;
        else probs[2] = 1.0f;
    }
};

// Layer table of Connect4Net (study-connect4/src/policies.rs:20-24)
static const int MLP_DIMS[6] = {63, 128, 96, 64, 48, 12};

// study-connect4/src/policies.rs:28-59 on weights in blob order l_1.weight, l_1.bias, ... l_5.bias
struct Connect4Net : Policy<Connect4> {
    const float* blob;
    bool libm;
    explicit Connect4Net(const float* weights, bool use_libm = false) : blob(weights), libm(use_libm) {}
    void forward(const float x0[63], float out12[12]) const {
        float a[128], b[128];
        const float* in = x0;
        float* bufs[2] = {a, b};
        const float* w = blob;
        for (int l = 0; l < 5; ++l) {
            int I = MLP_DIMS[l], O = MLP_DIMS[l + 1];
            const float* W = w;
            const float* B = w + (size_t)I * O;
            float* out = (l == 4) ? out12 : bufs[l & 1];
            for (int o = 0; o < O; ++o) out[o] = B[o];
            for (int i = 0; i < I; ++i)
                for (int o = 0; o < O; ++o) out[o] += in[i] * W[(size_t)o * I + i];
            if (l < 4)
                for (int o = 0; o < O; ++o) out[o] = out[o] > 0.0f ? out[o] : 0.0f; // relu
            in = out;
            w += (size_t)I * O + O;
        }
    }
    void eval(const Connect4& g, float logits[9], float probs[3]) override {
        float x[63], y[12];
        g.features(x);
        forward(x, y);
        for (int i = 0; i < 9; ++i) logits[i] = y[i];
        // value.softmax(-1): stable softmax of the last three outputs
        float m = std::fmax(y[9], std::fmax(y[10], y[11]));
        float e[3], t = 0.0f;
        for (int i = 0; i < 3; ++i) {
            e[i] = f_exp(y[9 + i] - m, libm);
            t += e[i];
        }
        for (int i = 0; i < 3; ++i) probs[i] = e[i] / t;
    }
};

// policies/cache.rs:19-32 (HashMap keyed by the game; Connect4 hashes its two bitboards)
template <class G>
struct PolicyWithCache : Policy<G> {
    struct Key {
        uint64_t lo, hi;
        bool operator==(const Key& o) const { return lo == o.lo && hi == o.hi; }
    };
    struct KeyHash {
        size_t operator()(const Key& k) const {
            uint64_t h = k.lo * 0x9e3779b97f4a7c15ull;
            h ^= (k.hi + 0x7f4a7c159e3779b9ull) * 0xc2b2ae3d27d4eb4full;
            return (size_t)(h ^ (h >> 29));
        }
    };
    struct Val { float logits[9]; float probs[3]; };
    Policy<G>* inner;
    std::unordered_map<Key, Val, KeyHash> cache;
    uint64_t hits = 0, misses = 0;
    PolicyWithCache(size_t capacity, Policy<G>* p) : inner(p) { cache.reserve(capacity); }
    void eval(const G& g, float logits[9], float probs[3]) override {
        Key k{g.key_lo(), g.key_hi()};
        auto it = cache.find(k);
        if (it != cache.end()) {
            ++hits;
            for (int i = 0; i < 9; ++i) logits[i] = it->second.logits[i];
            for (int i = 0; i < 3; ++i) probs[i] = it->second.probs[i];
            return;
        }
        ++misses;
        inner->eval(g, logits, probs);
        Val v;
        for (int i = 0; i < 9; ++i) v.logits[i] = logits[i];
        for (int i = 0; i < 3; ++i) v.probs[i] = probs[i];
        cache.emplace(k, v);
    }
};

// A Policy whose outputs come from outside (used to feed the oracle the GPU's leaf outputs).
typedef void (*orc_eval_fn)(void* ctx, uint64_t my_bb, uint64_t op_bb, float* logits9, float* probs3);
struct CallbackPolicy : Policy<Connect4> {
    orc_eval_fn fn;
    void* ctx;
    CallbackPolicy(orc_eval_fn f, void* c) : fn(f), ctx(c) {}
    void eval(const Connect4& g, float logits[9], float probs[3]) override { fn(ctx, g.my_bb, g.op_bb, logits, probs); }
};

} // namespace orc
