// mlp_tc.cuh — Connect4Net forward on the 5th-generation tensor cores (tcgen05 + TMEM), one
// CTA-resident batch of up to 128 leaves per call.
//
// Replaces study-connect4/src/policies.rs:28-59 (five nn::Linear + ReLU through libtorch, batch 1
// per leaf) with a chain of five UMMA GEMMs per batch:
//     D[128 x N] (fp32, TMEM) = A[128 x K] (fp16, smem, K-major) * W[N x K]^T (fp16, smem, K-major)
// with (K, N) = (80,128) (128,96) (96,64) (64,48) (48,16).  PyTorch keeps nn.Linear weights as
// [out][in] row-major, which IS the K-major B operand, so no transpose is needed.  Layer 0's K axis is
// permuted: feature (row, col) of Game::features (index row*9+col) sits at k = col*8 + row, i.e. one
// 16-byte chunk of 8 fp16 per board column (7 cells + a zero), chunks 0..8, chunk 9 all zero; the
// weight image carries the same permutation, so the product is unchanged and a column's features
// can be written with one 16-byte store (mlp_team.cuh looks them up in a 255-entry table).
//
//  * weights: converted once (syn_engine_set_weights) to an fp16 image in exactly the shared-memory
//    layout the UMMA descriptors expect, then brought into shared memory with ONE bulk-async copy
//    (cp.async.bulk, the TMA engine; UBLKCP in SASS) per CTA per kernel and kept resident.
//  * activations never leave the SM: features are synthesised from the 16-byte bitboards straight
//    into the A tile; each layer's accumulator is read back from TMEM (tcgen05.ld), gets bias + ReLU
//    in registers, is rounded to fp16 and written as the next layer's A tile.
//  * numerics: fp16 operands (11-bit significand, the same as TF32), fp32 accumulate.  Checked in
//    tests against the fp32 oracle forward to 1e-3 abs/rel (BASELINE.json north_star).
//
// Shared-memory operand layout (no swizzle, "INTERLEAVE" canonical K-major layout, cf. CUTLASS
// cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::K>): an R-row, K-column fp16 operand is
// stored as K/8 column chunks of [R rows][8 elements = 16 B]:
//     offset(r, k) = (k / 8) * (R * 16) + r * 16 + (k % 8) * 2      bytes
// so a core matrix (8 rows x 16 B) is 128 contiguous bytes, SBO (next 8-row group) = 128 B and
// LBO (next 8-column chunk) = R * 16 B.  One tcgen05.mma consumes K = 16, i.e. two chunks.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mlptc {

constexpr int M_TILE = 128;
constexpr int NL = 5;
__host__ __device__ constexpr int layer_k(int l) { return l == 0 ? 80 : (l == 1 ? 128 : (l == 2 ? 96 : (l == 3 ? 64 : 48))); }
__host__ __device__ constexpr int layer_n(int l) { return l == 0 ? 128 : (l == 1 ? 96 : (l == 2 ? 64 : (l == 3 ? 48 : 16))); }
__host__ __device__ constexpr int layer_in(int l) { return l == 0 ? 63 : layer_k(l); }   // real fan-in
// K index of feature i = row*9+col in layer 0's permuted K axis
__host__ __device__ constexpr int kperm(int i) { return (i % 9) * 8 + (i / 9); }
__host__ __device__ constexpr int layer_out(int l) { return l == 4 ? 12 : layer_n(l); }  // real fan-out
__host__ __device__ constexpr int w_bytes(int l) { return layer_k(l) * layer_n(l) * 2; }
// Byte offset of layer l's weights in the image.  Written WITHOUT recursion: a recursive constexpr function
// called with the (unrolled) loop variable is not a constant expression, and nvcc emitted real recursive calls
// with local-memory stack frames for it inside the forward pass (8 % of all stall samples in ncu).
__host__ __device__ constexpr int w_off(int l) {
    return l == 0 ? 0 : (l == 1 ? w_bytes(0) : (l == 2 ? w_bytes(0) + w_bytes(1) : (l == 3 ? w_bytes(0) + w_bytes(1) + w_bytes(2)
                                                                                          : w_bytes(0) + w_bytes(1) + w_bytes(2) + w_bytes(3))));
}
constexpr int W_TOTAL = w_off(4) + w_bytes(4);          // 65024 B of fp16 weights
constexpr int BIAS_OFF = W_TOTAL;                        // fp32 biases, padded to N
__host__ __device__ constexpr int b_off(int l) { // float offset of layer l's bias; no recursion (see w_off)
    return l == 0 ? 0 : (l == 1 ? layer_n(0) : (l == 2 ? layer_n(0) + layer_n(1) : (l == 3 ? layer_n(0) + layer_n(1) + layer_n(2)
                                                                                          : layer_n(0) + layer_n(1) + layer_n(2) + layer_n(3))));
}
static_assert(w_off(4) == 63488 && b_off(4) == 336, "layer table");
constexpr int BIAS_FLOATS = b_off(4) + layer_n(4);       // 352
constexpr int IMG_BYTES = W_TOTAL + BIAS_FLOATS * 4;     // 66432 B, multiple of 16
static_assert(IMG_BYTES % 16 == 0, "bulk copy size must be a multiple of 16 bytes");
// ONE activation tile, rewritten in place: a layer's epilogue starts after all of the layer's MMAs have completed
// (mbarrier), so the rows of the next layer's A operand can overwrite the rows just consumed.  32 KB instead of two
// tiles (56 KB) is what lets two CTAs share an SM (107 KB each with the weight image).
constexpr int A_BYTES = (128 / 8) * M_TILE * 16;         // K up to 128 -> 32768 B
constexpr int TMEM_COLS = 128;

// blob offsets (floats) of l_k.weight / l_k.bias in the caller's weight blob
__host__ __device__ constexpr int blob_sz(int l) { return layer_in(l) * layer_out(l) + layer_out(l); }
__host__ __device__ constexpr int blob_w(int l) {
    return l == 0 ? 0 : (l == 1 ? blob_sz(0) : (l == 2 ? blob_sz(0) + blob_sz(1) : (l == 3 ? blob_sz(0) + blob_sz(1) + blob_sz(2)
                                                                                          : blob_sz(0) + blob_sz(1) + blob_sz(2) + blob_sz(3))));
}
static_assert(blob_w(4) + blob_sz(4) == 30492, "Connect4Net has 30492 parameters");

// ---- one-time conversion of the fp32 blob to the shared-memory image (global memory)
__global__ void build_weight_image(const float* __restrict__ blob, uint8_t* __restrict__ img) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = layer_k(l), N = layer_n(l), I = layer_in(l), O = layer_out(l);
        const float* W = blob + blob_w(l);
        const float* B = W + I * O;
        __half* dst = reinterpret_cast<__half*>(img + w_off(l));
        for (int e = tid; e < K * N; e += nt) {
            int n = e / K, k = e - n * K;
            int src = k; // layer 0: k = col*8 + row  <-  blob column row*9 + col
            if (l == 0) src = ((k & 7) < 7 && (k >> 3) < 9) ? (k & 7) * 9 + (k >> 3) : I;
            float v = (n < O && src < I) ? W[n * I + src] : 0.0f;
            dst[(k / 8) * (N * 8) + n * 8 + (k % 8)] = __float2half_rn(v);
        }
        float* bd = reinterpret_cast<float*>(img + BIAS_OFF) + b_off(l);
        for (int e = tid; e < N; e += nt) bd[e] = e < O ? B[e] : 0.0f;
    }
}

// ---- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    // the suspend-time hint lets the warp sleep in hardware until the phase completes instead of re-issuing the
    // try_wait / branch / yield triple (14 % of all issued instructions in ncu before the hint)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    return ok != 0;
}
// (a __nanosleep between failed polls was measured and bought nothing: profiles/r1_mbar_sleep_ab.txt)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) { // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) { // one full warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format F16 (0),
// a/b K-major (0), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M_TILE >> 4) << 24); }

// Shared-memory block of the tensor-core MLP.
struct __align__(128) Smem {
    uint8_t img[IMG_BYTES];   // fp16 weights of the five layers + fp32 biases (the bulk-copied image)
    uint8_t a[A_BYTES];       // the A tile: features first, then every layer's activations in place
    float y[M_TILE][16];      // final layer output per row: 9 policy logits, 3 value logits, 4 pad
    uint64_t bar_w;           // weights landed
    uint64_t bar_mma;         // a layer's MMAs completed
    uint32_t tmem_base;
    uint32_t pad;
};

// tile row of leaf slot `i` (i < 128): spread over the four 32-lane TMEM quarters so that every
// epilogue warp has rows to work on
__device__ __forceinline__ int row_of_slot(int i) { return (i & 3) * 32 + (i >> 2); }

// Byte offset of element (row, k) in an A tile.
__device__ __forceinline__ uint32_t a_off(int row, int k) { return (uint32_t)((k >> 3) * (M_TILE * 16) + row * 16 + (k & 7) * 2); }

// Prologue: all threads call.  NW warps per CTA.
__device__ __forceinline__ void setup(Smem& s, const uint8_t* __restrict__ weight_image) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&s.bar_w, 1);
        mbar_init(&s.bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&s.tmem_base, TMEM_COLS);
    // zero the activation tiles once: rows that never carry a leaf must not hold NaN patterns that
    // could leak (they cannot: rows are independent) — zeroing just keeps the tile deterministic
    for (int i = threadIdx.x; i < A_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(s.a)[i] = make_uint4(0u, 0u, 0u, 0u);
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

__device__ __forceinline__ void teardown(Smem& s) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) tmem_dealloc(s.tmem_base, TMEM_COLS);
}

// Forward pass over the rows currently in s.a (fp16 features, K = 80, permuted: kperm()).  All NW warps call; every
// thread must have finished writing its part of the tile (generic-proxy stores) before the call — the
// function issues the proxy fence and the CTA barrier itself.  `phase` is the running parity of
// bar_mma and must be kept by the caller across calls (start at 0).  nrows_per_quarter = number
// of valid rows in each 32-row quarter (valid slots / 4, rounded up).
template <int NW>
__device__ __forceinline__ void forward(Smem& s, uint32_t& phase, int nrows_per_quarter) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quarter = warp & 3;
    const uint32_t tmem = s.tmem_base;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = layer_k(l), N = layer_n(l);
        uint8_t* a_in = s.a;
        uint8_t* a_out = s.a;
        if (threadIdx.x == 0) {
            tc_fence_after();
            const uint32_t a_base = smem_u32(a_in), b_base = smem_u32(s.img + w_off(l));
#pragma unroll
            for (int kk = 0; kk < K / 16; ++kk) {
                uint64_t ad = make_desc(a_base + kk * 2 * (M_TILE * 16), M_TILE * 16, 128);
                uint64_t bd = make_desc(b_base + kk * 2 * (N * 16), N * 16, 128);
                umma_f16(tmem, ad, bd, make_idesc(N), kk > 0 ? 1u : 0u);
            }
            umma_commit(&s.bar_mma);
        }
        mbar_wait(&s.bar_mma, phase);
        phase ^= 1u;
        tc_fence_after();
        // epilogue: 8-column slices; warp w takes TMEM quarter w % 4 and slices w/4, w/4 + NW/4, ...
        const float* bias = reinterpret_cast<const float*>(s.img + BIAS_OFF) + b_off(l);
        const int row = quarter * 32 + lane;
        const bool valid = lane < nrows_per_quarter;
#pragma unroll
        for (int sl = warp >> 2; sl < N / 8; sl += NW / 4) {
            uint32_t r[8];
            tmem_ld8(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sl * 8), r);
            tmem_ld_wait();
            if (valid) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) + bias[sl * 8 + j];
                if (l < NL - 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fminf(fmaxf(v[j], 0.0f), 65504.0f); // ReLU, clamp to fp16 range
                    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
                    __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
                    uint4 pk = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                          *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
                    *reinterpret_cast<uint4*>(a_out + a_off(row, sl * 8)) = pk;
                } else {
                    float4* yo = reinterpret_cast<float4*>(&s.y[row][sl * 8]);
                    yo[0] = make_float4(v[0], v[1], v[2], v[3]);
                    yo[1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
    }
}

} // namespace mlptc
