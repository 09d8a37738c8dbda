// oracle/rng.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the random streams the reference's hot path consumes.  The algorithms live
// in crates that are NOT vendored in /root/reference (Cargo.lock is git-ignored):
//   rand = "0.8.3"        (synthesis/Cargo.toml:12)  -> StdRng, gen_range, WeightedIndex, Uniform
//   rand_chacha 0.3.x     (transitive)               -> ChaCha12 core behind StdRng
//   rand_core 0.6.x       (transitive)               -> SeedableRng::seed_from_u64, BlockRng
//   rand_distr = "0.4.0"  (synthesis/Cargo.toml:20)  -> Dirichlet, Normal (see noise.hpp)
// Call sites in the reference: StdRng::seed_from_u64 (alpha_zero.rs:189, evaluator.rs:172,207,
// mcts.rs:692), gen_range(0..n as u8) (policies/rollout.rs:16, alpha_zero.rs:281),
// WeightedIndex::new(pi).sample (alpha_zero.rs:286-287).
//
// Pinning: StdRng + seed_from_u64 + gen_range<u8> are pinned end-to-end by the reference's own
// constants nodes.len() == 311 / 69 / 1533 (mcts.rs:732,781,830), reproduced by
// oracle/selftest.cpp in `legacy` tree mode, and by the ChaCha20 all-zero-key block from RFC 7539.
// Uniform<f32>/WeightedIndex are restated from rand 0.8's published algorithm; no reference test
// fixes one of their values ("parity unpinned").
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

static inline uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

// ChaCha block function with a configurable number of rounds (12 for StdRng in rand 0.8).
// State layout: 4 constants, 8 key words, 64-bit block counter in words 12-13, 64-bit stream id in
// words 14-15 (always 0 for seed_from_u64).
static inline void chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                       key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                       (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t x[16];
    std::memcpy(x, in, sizeof(x));
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    };
    for (int r = 0; r < rounds; r += 2) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
}

// rand 0.8 `StdRng` = ChaCha12Rng wrapped in BlockRng: the output is the concatenation of
// consecutive ChaCha12 blocks, consumed as little-endian u32 words in order.
struct StdRng {
    uint32_t key[8];
    uint64_t counter = 0; // next block to generate
    uint32_t buf[16];
    int idx = 16;
    uint64_t words_drawn = 0;

    // rand_core 0.6 SeedableRng::seed_from_u64: a PCG32 stream fills the 32-byte seed.
    static StdRng seed_from_u64(uint64_t state) {
        StdRng r;
        const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
        for (int i = 0; i < 8; ++i) {
            state = state * MUL + INC;
            uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            r.key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        }
        return r;
    }
    uint32_t next_u32() {
        if (idx == 16) {
            chacha_block(key, counter++, 12, buf);
            idx = 0;
        }
        ++words_drawn;
        return buf[idx++];
    }
    // rand 0.8 UniformInt<u8>::sample_single via gen_range(0..n): u8 is widened to u32, the
    // rejection zone is computed with a modulus (types <= 16 bit), widening multiply picks hi.
    uint32_t gen_range_u8(uint32_t n) {
        const uint32_t umax = 0xffffffffu;
        uint32_t ints_to_reject = (umax - n + 1u) % n;
        uint32_t zone = umax - ints_to_reject;
        for (;;) {
            uint32_t v = next_u32();
            uint64_t m = (uint64_t)v * (uint64_t)n;
            uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
            if (lo <= zone) return hi;
        }
    }
    // rand 0.8 Standard f32 in [0,1) as used inside UniformFloat::sample: 23 mantissa bits, [1,2)-1.
    float next_f32_01() {
        uint32_t v = next_u32();
        uint32_t bits = (v >> 9) | 0x3f800000u;
        float f;
        std::memcpy(&f, &bits, 4);
        return f - 1.0f;
    }
};

// rand 0.8 UniformFloat<f32>::new(low, high) + sample.
struct UniformF32 {
    float low, scale;
    UniformF32(float lo, float hi) : low(lo) {
        const float max_rand = 1.0f - 1.1920929e-7f; // (u32::MAX >> 9) as [1,2) float minus 1 = 1 - 2^-23
        scale = hi - lo;
        for (;;) {
            float top = scale * max_rand + lo;
            if (!(top >= hi)) break;
            // next representable float below scale
            uint32_t b;
            std::memcpy(&b, &scale, 4);
            b -= 1;
            std::memcpy(&scale, &b, 4);
        }
    }
    float sample(StdRng& r) const {
        float v01 = r.next_f32_01();
        return v01 * scale + low;
    }
};

// rand 0.8 WeightedIndex<f32>: cumulative sums of all but the last weight, Uniform(0,total),
// index = number of cumulative weights <= draw.  Returns -1 where the reference would
// `.unwrap()`-panic (no weights, negative/NaN weight, zero total).
static inline int weighted_index_sample(const float* w, int n, StdRng& r) {
    if (n <= 0) return -1;
    std::vector<float> cum;
    float total = w[0];
    if (!(total >= 0.0f)) return -1;
    for (int i = 1; i < n; ++i) {
        if (!(w[i] >= 0.0f)) return -1;
        cum.push_back(total);
        total += w[i];
    }
    if (!(total > 0.0f)) return -1;
    UniformF32 u(0.0f, total);
    float x = u.sample(r);
    int idx = 0;
    for (float c : cum)
        if (c <= x) ++idx;
        else break;
    return idx;
}

} // namespace orc
