// rng.cuh — the reference's random streams on the device.
//
// rand 0.8 `StdRng` = ChaCha12 (rand_chacha 0.3), seeded by rand_core 0.6 `seed_from_u64`
// (PCG32 expansion), consumed as consecutive little-endian u32 words of consecutive blocks.
// Call sites replaced: policies/rollout.rs:16 (one u32 per rollout ply), alpha_zero.rs:281,
// 286-287 (action sampling).  The CPU checker restates the same streams (tests compare them word for word).
#pragma once
#include <stdint.h>

namespace rng {

__host__ __device__ __forceinline__ uint32_t rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

// rand_core 0.6 SeedableRng::seed_from_u64
__host__ __device__ inline void seed_key(uint64_t state, uint32_t key[8]) {
    const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        state = state * MUL + INC;
        uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
}

#define SYN_QR(a, b, c, d)                  \
    a += b; d = rng::rotl(d ^ a, 16);       \
    c += d; b = rng::rotl(b ^ c, 12);       \
    a += b; d = rng::rotl(d ^ a, 8);        \
    c += d; b = rng::rotl(b ^ c, 7);

// One whole ChaCha12 block by one thread (cold paths: action sampling, noise streams).
__device__ __noinline__ void chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t x0 = 0x61707865u, x1 = 0x3320646eu, x2 = 0x79622d32u, x3 = 0x6b206574u;
    uint32_t x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3], x8 = key[4], x9 = key[5], x10 = key[6], x11 = key[7];
    uint32_t x12 = (uint32_t)counter, x13 = (uint32_t)(counter >> 32), x14 = 0u, x15 = 0u;
#pragma unroll 1
    for (int r = 0; r < 6; ++r) {
        SYN_QR(x0, x4, x8, x12) SYN_QR(x1, x5, x9, x13) SYN_QR(x2, x6, x10, x14) SYN_QR(x3, x7, x11, x15)
        SYN_QR(x0, x5, x10, x15) SYN_QR(x1, x6, x11, x12) SYN_QR(x2, x7, x8, x13) SYN_QR(x3, x4, x9, x14)
    }
    out[0] = x0 + 0x61707865u; out[1] = x1 + 0x3320646eu; out[2] = x2 + 0x79622d32u; out[3] = x3 + 0x6b206574u;
    out[4] = x4 + key[0]; out[5] = x5 + key[1]; out[6] = x6 + key[2]; out[7] = x7 + key[3];
    out[8] = x8 + key[4]; out[9] = x9 + key[5]; out[10] = x10 + key[6]; out[11] = x11 + key[7];
    out[12] = x12 + (uint32_t)counter; out[13] = x13 + (uint32_t)(counter >> 32); out[14] = x14; out[15] = x15;
}

// A sequential stream for cold paths, used by one thread.  R-concept of include/syn_sampling.h.
struct Stream {
    uint32_t key[8];
    uint32_t buf[16];
    uint64_t pos;      // next word index in the stream
    uint64_t have_blk; // block currently in buf (or ~0)
    __device__ void init(uint64_t seed, uint64_t start_pos) {
        seed_key(seed, key);
        pos = start_pos;
        have_blk = ~0ull;
    }
    __device__ uint32_t next_u32() {
        uint64_t blk = pos >> 4;
        if (blk != have_blk) {
            chacha12_block(key, blk, buf);
            have_blk = blk;
        }
        return buf[(pos++) & 15];
    }
    // rand 0.8 gen_range(0..n) for u8 (UniformInt::sample_single, u32-widened, modulus zone)
    __device__ uint32_t gen_range_u8(uint32_t n) {
        uint32_t zone = 0xffffffffu - ((0xffffffffu - n + 1u) % n);
        for (;;) {
            uint32_t v = next_u32();
            uint64_t m = (uint64_t)v * (uint64_t)n;
            if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
        }
    }
    // the same for n in 1..9 (a Connect4 move count) without the modulo: 2^32 mod n comes from a packed table
    __device__ uint32_t gen_range_1to9(uint32_t n) {
        const uint32_t zone = 0xffffffffu - (uint32_t)((0x4044101000ull >> (4u * n)) & 15ull);
        for (;;) {
            uint32_t v = next_u32();
            uint64_t m = (uint64_t)v * (uint64_t)n;
            if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
        }
    }
    __device__ float next_f32_01() { return __uint_as_float((next_u32() >> 9) | 0x3f800000u) - 1.0f; }
};

} // namespace rng
