// mlp_team.cuh — Connect4Net forward for a TEAM of 128 threads = 128 leaves = one UMMA M-tile.
//
// Replaces study-connect4/src/policies.rs:28-59 (five nn::Linear + ReLU through libtorch, batch 1
// per leaf).  Same GEMM chain, operand layout and numerics as mlp_tc.cuh (fp16 operands, fp32
// accumulate in TMEM, bias + ReLU in registers), re-cut for the thread-per-game kernels:
//
//   * thread r of the team owns tile row r end to end: it writes its leaf's 64 fp16 features into
//     the A tile, and after every layer reads ITS row of the accumulator back from TMEM lane r
//     (tcgen05.ld 32x32b: warp w of the team covers lanes 32w..32w+31), applies bias + ReLU, and
//     writes the row of the next layer's A tile.  The final 12 outputs arrive in the registers of
//     the thread whose tree needs them — no staging of results through shared memory.
//   * one A tile per team: a layer's MMAs have completed (mbarrier) before any thread writes the
//     next activations, so layer l+1's A overwrites layer l's in place.
//   * several teams per CTA share ONE resident fp16 weight image (bulk-async copy, once per
//     kernel) and run their chains independently: team barriers are named barriers over 128
//     threads; a forward runs in one of SLOTS "MLP slots" = {A tile, 128 TMEM columns, mbarrier}.
//     With SLOTS == TEAMS every team owns a slot; with more teams than slots (8 teams, 4 slots:
//     1024 threads per SM at 64 registers) a team takes any free slot for the ~15 % of a round
//     it spends in the forward and gives it back, so shared memory and TMEM hold four tiles
//     while twice as many warps hide the tree phase's memory latency.
#pragma once
#include "mlp_tc.cuh"

namespace mlpteam {

using namespace mlptc;

constexpr int A_BYTES = (128 / 8) * M_TILE * 16; // K up to 128 -> 32768 B per team
constexpr int TEAM = 128;

template <int TEAMS, int SLOTS>
struct __align__(128) Smem {
    uint8_t img[IMG_BYTES];      // fp16 weights of the five layers + fp32 biases (bulk-copied image)
    uint8_t a[SLOTS][A_BYTES];   // one activation tile per MLP slot
    uint4 col_lut[256];          // fp16 features of one board column by (stones, owners): see write_features
    uint64_t bar_w;              // weights landed
    uint64_t bar_mma[SLOTS];     // a slot's layer completed
    uint32_t slot_busy[SLOTS];   // 0 free / 1 taken (only used when SLOTS < TEAMS)
    uint32_t slot_phase[SLOTS];  // running parity of bar_mma[slot], handed from owner to owner
    uint32_t team_slot[TEAMS];   // the slot a team's leader just took (broadcast to the team)
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void team_sync(int team) { // named barrier 1 + team over the team's 128 threads
    asm volatile("barrier.sync %0, %1;" ::"r"(team + 1), "n"(TEAM) : "memory");
}
__device__ __forceinline__ bool team_any(int team, bool p) {
    uint32_t out;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %2, 0;\n\tbarrier.red.or.pred q, %1, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(out) : "r"(team + 1), "r"((uint32_t)p), "n"(TEAM) : "memory");
    return out != 0u;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}

// {lo, hi} -> fp16x2 with ReLU and saturation to +-65504 (F2FP.SATFINITE.RELU.F16.F32.PACK_AB)
__device__ __forceinline__ uint32_t cvt_relu_sat_f16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// Two IEEE fp32 additions in one instruction (FADD2): each lane of the pair is rounded to nearest
// exactly like a scalar add, so results are unchanged.
__device__ __forceinline__ void add2(float a0, float a1, float b0, float b1, float& r0, float& r1) {
    asm("{\n\t.reg .b64 a, b, r;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 r, a, b;\n\tmov.b64 {%0, %1}, r;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// Game::features (connect4.rs:237-258) of one position, fp16, written as row `r` of the team's A tile
// in the K-major UMMA layout with layer 0's permuted K axis (mlp_tc.cuh): chunk c = board column c =
// 7 cells bottom-up + a zero.  +1 mine, -1 theirs, +0.1 the next playable cell, -0.1 any other empty
// cell.  A column with h stones whose owners are the low h bits of my_col is entry
// (2^h - 1) + my_col = occ_col + my_col of a 255-entry table, so a column costs one 16-byte
// table read and one 16-byte store.
__device__ __forceinline__ uint4 col_lut_entry(int idx) {
    int h = 31 - __clz(idx + 1);
    uint32_t mine = (uint32_t)(idx + 1) - (1u << h);
    uint32_t hv[8];
#pragma unroll
    for (int row = 0; row < 8; ++row) {
        uint32_t v;
        if (row == 7) v = 0u;
        else if (row < h) v = ((mine >> row) & 1u) ? 0x3C00u : 0xBC00u; // fp16 +1 / -1
        else if (row == h) v = 0x2E66u;                                  // fp16 RN(+0.1)
        else v = 0xAE66u;                                                // fp16 RN(-0.1)
        hv[row] = v;
    }
    return make_uint4(hv[0] | (hv[1] << 16), hv[2] | (hv[3] << 16), hv[4] | (hv[5] << 16), hv[6] | (hv[7] << 16));
}

__device__ __forceinline__ void write_features(uint8_t* a_tile, const uint4* lut, int r, uint64_t my, uint64_t op) {
    const uint64_t occ = my | op;
#pragma unroll
    for (int col = 0; col < 9; ++col) {
        uint32_t oc = (uint32_t)(occ >> (7 * col)) & 0x7fu, mc = (uint32_t)(my >> (7 * col)) & 0x7fu;
        *reinterpret_cast<uint4*>(a_tile + col * (M_TILE * 16) + r * 16) = lut[oc + mc];
    }
    *reinterpret_cast<uint4*>(a_tile + 9 * (M_TILE * 16) + r * 16) = make_uint4(0u, 0u, 0u, 0u); // K 72..79
}

// Prologue: all threads of the CTA call.
template <int TEAMS, int SLOTS>
__device__ __forceinline__ void setup(Smem<TEAMS, SLOTS>& s, const uint8_t* __restrict__ weight_image) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&s.bar_w, 1);
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) { mbar_init(&s.bar_mma[t], 1); s.slot_busy[t] = 0u; s.slot_phase[t] = 0u; }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&s.tmem_base, 128 * SLOTS);
    for (int i = threadIdx.x; i < SLOTS * A_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(&s.a[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s.col_lut[i] = col_lut_entry(i < 255 ? i : 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&s.bar_w, IMG_BYTES);
        bulk_g2s(s.img, weight_image, IMG_BYTES, &s.bar_w);
    }
    mbar_wait(&s.bar_w, 0);
    fence_proxy_async();
    __syncthreads();
}

template <int TEAMS, int SLOTS>
__device__ __forceinline__ void teardown(Smem<TEAMS, SLOTS>& s) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(s.tmem_base, 128 * SLOTS);
}

// Takes an MLP slot for the team (all 128 threads call; returns the slot and the parity of its mbarrier).
template <int TEAMS, int SLOTS>
__device__ __forceinline__ int acquire_slot(Smem<TEAMS, SLOTS>& s, int team, int r, uint32_t& phase) {
    if (SLOTS == TEAMS) { phase = s.slot_phase[team]; return team; }
    if (r == 0) {
        int got = -1;
        for (int k = team; got < 0; ++k) {
            int cand = k % SLOTS;
            if (atomicCAS(&s.slot_busy[cand], 0u, 1u) == 0u) got = cand;
            else if (cand == SLOTS - 1) __nanosleep(64);
        }
        __threadfence_block();
        s.team_slot[team] = (uint32_t)got;
    }
    team_sync(team);
    int slot = (int)s.team_slot[team];
    phase = s.slot_phase[slot];
    return slot;
}

// Gives the slot back: every thread has finished reading the slot's TMEM columns and A tile.
template <int TEAMS, int SLOTS>
__device__ __forceinline__ void release_slot(Smem<TEAMS, SLOTS>& s, int team, int r, int slot, uint32_t phase) {
    if (SLOTS == TEAMS) { if (r == 0) s.slot_phase[slot] = phase; return; }
    tc_fence_before();
    team_sync(team);
    if (r == 0) {
        s.slot_phase[slot] = phase;
        __threadfence_block();
        atomicExch(&s.slot_busy[slot], 0u);
    }
}

// Hidden-layer epilogue for 16 accumulator columns of this thread's row: bias (two lanes per FADD2), then ReLU + saturate to
// the fp16 range + round + pack in ONE instruction, stored as the row's next-layer A operand (K chunks 2*c16, 2*c16+1).
__device__ __forceinline__ void epilogue16(const uint32_t (&v)[16], const float* bias, uint8_t* a_tile, int r, int c16) {
    uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 b4 = *reinterpret_cast<const float4*>(bias + c16 * 16 + 4 * j);
        float x0, x1, x2, x3;
        add2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), b4.x, b4.y, x0, x1);
        add2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]), b4.z, b4.w, x2, x3);
        h[2 * j] = cvt_relu_sat_f16x2(x0, x1);
        h[2 * j + 1] = cvt_relu_sat_f16x2(x2, x3);
    }
    *reinterpret_cast<uint4*>(a_tile + (2 * c16) * (M_TILE * 16) + r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_tile + (2 * c16 + 1) * (M_TILE * 16) + r * 16) = make_uint4(h[4], h[5], h[6], h[7]);
}

// epilogue16 with the biases in the kernel's parameter space (constant bank: an operand of the FADD, no load at all).
// The shared-memory form waits for four LDS.128 per 16 columns on the short scoreboard — 24 % of all stall samples of
// a kernel whose tree code waited less (profiles/r2_tpg3_t5_source_summary.txt).
__device__ __forceinline__ void epilogue16_cb(const uint32_t (&v)[16], const float* __restrict__ bias, uint8_t* a_tile, int r, int c16) {
    uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float x0 = __fadd_rn(__uint_as_float(v[2 * j]), bias[c16 * 16 + 2 * j]);
        const float x1 = __fadd_rn(__uint_as_float(v[2 * j + 1]), bias[c16 * 16 + 2 * j + 1]);
        h[j] = cvt_relu_sat_f16x2(x0, x1);
    }
    *reinterpret_cast<uint4*>(a_tile + (2 * c16) * (M_TILE * 16) + r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_tile + (2 * c16 + 1) * (M_TILE * 16) + r * 16) = make_uint4(h[4], h[5], h[6], h[7]);
}

// Forward pass of the team's 128 rows.  Every thread of the team calls with its row's features
// already in the slot's A tile (generic-proxy stores); `phase` is the running parity of the slot's
// mbarrier (acquire_slot / release_slot carry it).  On return y[0..8] are the row's policy logits and
// y[9..11] its value logits.
// `img` = the resident weight image to use (s.img, or a second image of the same layout elsewhere in shared memory).
// `cbias` = the image's 352 padded biases in the kernel's parameter space, or nullptr to read them from the image.
template <int TEAMS, int SLOTS, bool CB = false>
__device__ __forceinline__ void forward_img(Smem<TEAMS, SLOTS>& s, const uint8_t* img, int team, int slot, int r, uint32_t& phase, float (&y)[12],
                                            const float* __restrict__ cbias = nullptr) {
    uint8_t* a_tile = s.a[slot];
    const uint32_t tmem = s.tmem_base + (uint32_t)(slot * 128);           // the slot's 128 accumulator columns
    const uint32_t tlane = tmem + ((uint32_t)((r >> 5) * 32) << 16);       // this warp's 32 TMEM lanes
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = layer_k(l), N = layer_n(l);
        fence_proxy_async();  // this thread's A-tile stores -> visible to the tensor core's async proxy
        tc_fence_before();
        team_sync(team);
        if (r == 0) {
            tc_fence_after();
            const uint32_t a_base = smem_u32(a_tile), b_base = smem_u32(img + w_off(l));
#pragma unroll
            for (int kk = 0; kk < K / 16; ++kk) {
                uint64_t ad = make_desc(a_base + kk * 2 * (M_TILE * 16), M_TILE * 16, 128);
                uint64_t bd = make_desc(b_base + kk * 2 * (N * 16), N * 16, 128);
                umma_f16(tmem, ad, bd, make_idesc(N), kk > 0 ? 1u : 0u);
            }
            umma_commit(&s.bar_mma[slot]);
        }
        mbar_wait(&s.bar_mma[slot], phase);
        phase ^= 1u;
        tc_fence_after();
        const float* bias = (CB ? cbias : reinterpret_cast<const float*>(img + BIAS_OFF)) + b_off(l);
        // (two 16-column reads in flight per wait were measured: -0.4 %, profiles/r1_tmem_pair_ab.txt)
#pragma unroll
        for (int c16 = 0; c16 < N / 16; ++c16) {
            uint32_t v[16];
            tmem_ld16(tlane + (uint32_t)(c16 * 16), v);
            tmem_ld_wait();
            if (l < NL - 1) {
                if (CB) epilogue16_cb(v, bias, a_tile, r, c16);
                else epilogue16(v, bias, a_tile, r, c16);
            } else {
#pragma unroll
                for (int j = 0; j < 12; ++j) y[j] = __uint_as_float(v[j]) + bias[j];
            }
        }
    }
    // the next forward()'s first team_sync orders these TMEM reads before the next MMA overwrites D
    tc_fence_before();
}

template <int TEAMS, int SLOTS>
__device__ __forceinline__ void forward(Smem<TEAMS, SLOTS>& s, int team, int slot, int r, uint32_t& phase, float (&y)[12]) {
    forward_img<TEAMS, SLOTS>(s, s.img, team, slot, r, phase, y);
}
// The same with the biases read from the kernel's parameter space (KParams::mlp_bias).
template <int TEAMS, int SLOTS>
__device__ __forceinline__ void forward_cb(Smem<TEAMS, SLOTS>& s, const float* __restrict__ cbias, int team, int slot, int r, uint32_t& phase, float (&y)[12]) {
    forward_img<TEAMS, SLOTS, true>(s, s.img, team, slot, r, phase, y, cbias);
}

// A second resident image (evaluation matches between two different networks): bulk-copied behind the first on the
// same mbarrier.  All threads of the CTA call, after setup().
template <int TEAMS, int SLOTS>
__device__ __forceinline__ void load_second_image(Smem<TEAMS, SLOTS>& s, uint8_t* dst, const uint8_t* __restrict__ weight_image) {
    if (threadIdx.x == 0) {
        mbar_expect_tx(&s.bar_w, IMG_BYTES);
        bulk_g2s(dst, weight_image, IMG_BYTES, &s.bar_w);
    }
    mbar_wait(&s.bar_w, 1);
    fence_proxy_async();
    __syncthreads();
}

} // namespace mlpteam
