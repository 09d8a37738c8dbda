// mlp_split.cuh — Connect4Net forward for a TEAM of 128 threads at fp32-grade accuracy on the tcgen05 tensor cores:
// split-fp16 operands (x = hi + lo, three MMAs per K-step: hi*Wh + hi*Wl + lo*Wh), fp32 accumulate, and the A operand
// kept in TENSOR MEMORY between layers instead of shared memory.
//
// Replaces study-connect4/src/policies.rs:28-59 (five nn::Linear + ReLU through libtorch in fp32, batch 1 per leaf).
// The single-fp16 chain of mlp_team.cuh meets BASELINE.json's 1e-3 only at initialisation scale: with trained-size weights
// (logits ~10) its error is 7x the tolerance (a numpy model of both chains: max error / (1e-3 + 1e-3 |ref|) = 7.4 against
// 0.004 for this chain; tests/test_gpu_parity.py asserts the unrelaxed tolerance).  The tensor pipe was 8 % busy, so the
// extra MMAs are free; what it costs is the epilogue's split (6 more instructions per pair of activations).
//
//  * activations never touch shared memory.  Thread r of the team owns tile row r = TMEM lane r end to end: it writes its
//    leaf's features as A rows with tcgen05.st, and after every layer reads ITS accumulator row (tcgen05.ld), adds the bias
//    (a constant-bank operand: KParams::mlp_bias), applies ReLU, splits every value into hi (the top 11 significant bits, exact
//    in fp16) and lo (the fp16-rounded remainder) and stores them IN PLACE over the accumulator columns just read: 16 fp32
//    columns become 8 columns of hi pairs + 8 columns of lo pairs = the A operand of one K-step of the next layer.
//    tcgen05.mma takes A straight from tensor memory ([a_tmem] operand form), B = the resident weight images from shared memory.
//  * a slot = 256 TMEM columns: regions X (128) and Y (96) alternate as accumulator and operand,
//        A0 -> Y, D0 -> X | A1 = X, D1 -> Y | A2 = Y, D2 -> X | A3 = X, D3 -> Y | A4 = Y, D4 -> X
//    so two slots fill the SM's 512 columns; shared memory holds only the two weight images (Wh, Wl = fp16(W - Wh), 65 KB
//    each), the feature table and the path table: 128 KB less than mlp_team.cuh, which the L1 gets back.
//  * error: |y - y_fp32| <= ~2^-21 relative per product (the dropped lo*Wl term) + fp32 accumulation order.
#pragma once
#include "mlp_tc.cuh"

namespace mlps {

using namespace mlptc;

constexpr int TEAM = 128;
constexpr int SLOT_COLS = 256; // TMEM columns per slot: X = [0, 128), Y = [128, 224)
constexpr int Y_OFF = 128;
constexpr int IMG2_BYTES = 2 * W_TOTAL; // Wh image, then Wl image (same UMMA layout as mlptc's weights; biases live in KParams)

struct Bias { float b[BIAS_FLOATS]; }; // the padded biases as a kernel parameter (constant bank)

template <int TEAMS, int SLOTS>
struct __align__(128) Smem {
    uint8_t img[IMG2_BYTES];     // fp16 hi weights of the five layers, then fp16 lo weights
    uint4 lut_hi[256];           // fp16 hi features of one board column by (stones, owners)
    uint4 lut_lo[256];           // .. and their lo parts (+-0.1 is not an fp16 number)
    uint64_t bar_w;              // weights landed
    uint64_t bar_mma[SLOTS];     // a slot's layer completed
    uint32_t slot_busy[SLOTS];   // 0 free / 1 taken (only used when SLOTS < TEAMS)
    uint32_t slot_phase[SLOTS];  // running parity of bar_mma[slot], handed from owner to owner
    uint32_t team_slot[TEAMS];   // the slot a team's leader just took (broadcast to the team)
    uint32_t tmem_base;
    uint32_t pad;
};

// ---- the lo image: Wl = fp16(W - float(fp16(W))), same element order as mlptc::build_weight_image's Wh
__global__ void build_weight_image_lo(const float* __restrict__ blob, uint8_t* __restrict__ img_lo) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = layer_k(l), N = layer_n(l), I = layer_in(l), O = layer_out(l);
        const float* W = blob + blob_w(l);
        __half* dst = reinterpret_cast<__half*>(img_lo + w_off(l));
        for (int e = tid; e < K * N; e += nt) {
            int n = e / K, k = e - n * K;
            int src = k;
            if (l == 0) src = ((k & 7) < 7 && (k >> 3) < 9) ? (k & 7) * 9 + (k >> 3) : I;
            float v = (n < O && src < I) ? W[n * I + src] : 0.0f;
            float hi = __half2float(__float2half_rn(v));
            dst[(k / 8) * (N * 8) + n * 8 + (k % 8)] = __float2half_rn(__fsub_rn(v, hi));
        }
    }
}

// ---- PTX wrappers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand read from tensor memory (row i = lane i, two fp16 per 32-bit column)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
// {lo, hi} -> fp16x2 with ReLU and saturation (F2FP.SATFINITE.RELU.F16.F32.PACK_AB); element `lo` goes to the low half = the even k
__device__ __forceinline__ uint32_t cvt_relu_sat_f16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// x = hi + lo: hi = x with its significand cut to 11 bits (exactly an fp16 number in fp16's normal range), lo = the rest,
// both through ReLU: a negative x gives (0, 0) because truncation and remainder keep x's sign.
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi2, uint32_t& lo2) {
    const float h0 = __uint_as_float(__float_as_uint(x0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(x1) & 0xffffe000u);
    hi2 = cvt_relu_sat_f16x2(h0, h1);
    lo2 = cvt_relu_sat_f16x2(__fsub_rn(x0, h0), __fsub_rn(x1, h1));
}

__device__ __forceinline__ void team_sync(int team) { asm volatile("barrier.sync %0, %1;" ::"r"(team + 1), "n"(TEAM) : "memory"); }
__device__ __forceinline__ bool team_any(int team, bool p) {
    uint32_t out;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %2, 0;\n\tbarrier.red.or.pred q, %1, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(out) : "r"(team + 1), "r"((uint32_t)p), "n"(TEAM) : "memory");
    return out != 0u;
}

// fp16 hi / lo features of one board column (see mlp_team.cuh col_lut_entry): +1 mine, -1 theirs, +0.1 the next playable cell,
// -0.1 any other empty cell; 7 cells bottom-up + a zero.
__device__ __forceinline__ void col_lut_entries(int idx, uint4& hi, uint4& lo) {
    int h = 31 - __clz(idx + 1);
    uint32_t mine = (uint32_t)(idx + 1) - (1u << h);
    uint32_t hv[8], lv[8];
#pragma unroll
    for (int row = 0; row < 8; ++row) {
        float v;
        if (row == 7) v = 0.0f;
        else if (row < h) v = ((mine >> row) & 1u) ? 1.0f : -1.0f;
        else if (row == h) v = 0.1f;
        else v = -0.1f;
        const float hf = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        hv[row] = (uint32_t)__half_as_ushort(__float2half_rn(hf));
        lv[row] = (uint32_t)__half_as_ushort(__float2half_rn(__fsub_rn(v, hf)));
    }
    hi = make_uint4(hv[0] | (hv[1] << 16), hv[2] | (hv[3] << 16), hv[4] | (hv[5] << 16), hv[6] | (hv[7] << 16));
    lo = make_uint4(lv[0] | (lv[1] << 16), lv[2] | (lv[3] << 16), lv[4] | (lv[5] << 16), lv[6] | (lv[7] << 16));
}

// Prologue: all threads of the CTA call.  img_hi / img_lo = the two weight images in global memory (W_TOTAL bytes each).
template <int TEAMS, int SLOTS>
__device__ __forceinline__ void setup(Smem<TEAMS, SLOTS>& s, const uint8_t* __restrict__ img_hi, const uint8_t* __restrict__ img_lo) {
    static_assert(SLOTS * SLOT_COLS <= 512, "a slot takes 256 of the SM's 512 tensor-memory columns");
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&s.bar_w, 1);
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) { mbar_init(&s.bar_mma[t], 1); s.slot_busy[t] = 0u; s.slot_phase[t] = 0u; }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&s.tmem_base, SLOTS * SLOT_COLS);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) col_lut_entries(i < 255 ? i : 0, s.lut_hi[i], s.lut_lo[i]);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&s.bar_w, IMG2_BYTES);
        bulk_g2s(s.img, img_hi, W_TOTAL, &s.bar_w);
        bulk_g2s(s.img + W_TOTAL, img_lo, W_TOTAL, &s.bar_w);
    }
    mbar_wait(&s.bar_w, 0);
    fence_proxy_async();
    __syncthreads();
}

template <int TEAMS, int SLOTS>
__device__ __forceinline__ void teardown(Smem<TEAMS, SLOTS>& s) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(s.tmem_base, SLOTS * SLOT_COLS);
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

// This thread's TMEM lane window and the slot's first column.
template <int TEAMS, int SLOTS>
__device__ __forceinline__ uint32_t lane_addr(const Smem<TEAMS, SLOTS>& s, int slot, int r) {
    return s.tmem_base + (uint32_t)(slot * SLOT_COLS) + ((uint32_t)((r >> 5) * 32) << 16);
}

// Game::features (connect4.rs:237-258) of one position as row r of layer 0's A operand, hi and lo, in region Y of the slot:
// K-step j (16 features = board columns 2j and 2j+1, see mlp_tc.cuh for the permuted K axis) = 8 columns of hi pairs + 8 of lo pairs.
// `live` = false writes zeros (a row without a leaf this round must not feed stale NaN patterns to the MMA: rows are independent,
// zeros just keep the tile deterministic).
template <int TEAMS, int SLOTS>
__device__ __forceinline__ void write_features(const Smem<TEAMS, SLOTS>& s, int slot, int r, uint64_t my, uint64_t op, bool live) {
    const uint64_t occ = my | op;
    const uint32_t base = lane_addr(s, slot, r) + (uint32_t)Y_OFF;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        uint4 h0 = make_uint4(0u, 0u, 0u, 0u), h1 = h0, l0 = h0, l1 = h0;
        if (live) {
            const int c0 = 2 * j, c1 = 2 * j + 1;
            const uint32_t i0 = ((uint32_t)(occ >> (7 * c0)) & 0x7fu) + ((uint32_t)(my >> (7 * c0)) & 0x7fu);
            h0 = s.lut_hi[i0]; l0 = s.lut_lo[i0];
            if (c1 < 9) {
                const uint32_t i1 = ((uint32_t)(occ >> (7 * c1)) & 0x7fu) + ((uint32_t)(my >> (7 * c1)) & 0x7fu);
                h1 = s.lut_hi[i1]; l1 = s.lut_lo[i1];
            }
        }
        const uint32_t v[16] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        tmem_st16(base + (uint32_t)(16 * j), v);
    }
}

// Forward pass of the team's 128 rows.  Every thread of the team calls with its row's features already stored (write_features);
// `phase` is the running parity of the slot's mbarrier; cbias = the 352 padded biases in the kernel's parameter space.  On
// return y[0..8] are the row's policy logits and y[9..11] its value logits.
template <int TEAMS, int SLOTS>
__device__ __forceinline__ void forward(Smem<TEAMS, SLOTS>& s, const float* __restrict__ cbias, int team, int slot, int r, uint32_t& phase, float (&y)[12]) {
    const uint32_t tslot = s.tmem_base + (uint32_t)(slot * SLOT_COLS); // MMA operands: lane field 0
    const uint32_t tlane = lane_addr(s, slot, r);                        // this warp's 32 lanes
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = layer_k(l), N = layer_n(l);
        const uint32_t a_off = (l & 1) ? 0u : (uint32_t)Y_OFF;   // operand region: Y for even layers, X for odd
        const uint32_t d_off = (l & 1) ? (uint32_t)Y_OFF : 0u;   // accumulator region: the other one
        tmem_st_wait();       // this thread's A rows are in tensor memory
        tc_fence_before();
        team_sync(team);
        if (r == 0) {
            tc_fence_after();
            const uint32_t b_hi = smem_u32(s.img + w_off(l)), b_lo = smem_u32(s.img + W_TOTAL + w_off(l));
#pragma unroll
            for (int kk = 0; kk < K / 16; ++kk) {
                const uint64_t bh = make_desc(b_hi + kk * 2 * (N * 16), N * 16, 128), bl = make_desc(b_lo + kk * 2 * (N * 16), N * 16, 128);
                const uint32_t a_hi = tslot + a_off + (uint32_t)(16 * kk), a_lo = a_hi + 8u;
                umma_f16_ts(tslot + d_off, a_hi, bh, make_idesc(N), kk > 0 ? 1u : 0u);
                umma_f16_ts(tslot + d_off, a_hi, bl, make_idesc(N), 1u);
                umma_f16_ts(tslot + d_off, a_lo, bh, make_idesc(N), 1u);
            }
            umma_commit(&s.bar_mma[slot]);
        }
        mbar_wait(&s.bar_mma[slot], phase);
        phase ^= 1u;
        tc_fence_after();
        const float* bias = cbias + b_off(l);
        // the read of chunk c+1 travels while chunk c is split and stored back (a wait::ld covers every read issued before it)
        uint32_t v[2][16];
        tmem_ld16(tlane + d_off, v[0]);
#pragma unroll
        for (int c16 = 0; c16 < N / 16; ++c16) {
            tmem_ld_wait();
            if (c16 + 1 < N / 16) tmem_ld16(tlane + d_off + (uint32_t)((c16 + 1) * 16), v[(c16 + 1) & 1]);
            const uint32_t (&cur)[16] = v[c16 & 1];
            if (l < NL - 1) {
                uint32_t o[16]; // in place: 8 columns of hi pairs, 8 of lo pairs = K-step c16 of the next layer's operand
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x0 = __fadd_rn(__uint_as_float(cur[2 * j]), bias[c16 * 16 + 2 * j]);
                    const float x1 = __fadd_rn(__uint_as_float(cur[2 * j + 1]), bias[c16 * 16 + 2 * j + 1]);
                    split_pair(x0, x1, o[j], o[8 + j]);
                }
                tmem_st16(tlane + d_off + (uint32_t)(c16 * 16), o);
            } else {
#pragma unroll
                for (int j = 0; j < 12; ++j) y[j] = __fadd_rn(__uint_as_float(cur[j]), bias[j]);
            }
        }
    }
    // the next forward's first team_sync orders these TMEM accesses before its MMAs
    tc_fence_before();
}

} // namespace mlps
