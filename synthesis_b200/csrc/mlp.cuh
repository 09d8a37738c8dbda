// mlp.cuh — Connect4Net forward for a CTA-resident batch of leaves (fp32 CUDA-core version).
//
// Replaces study-connect4/src/policies.rs:28-59 (five nn::Linear + ReLU, libtorch CPU, batch 1)
// for the leaves of all games a CTA owns.  Weights stay resident in shared memory for the
// lifetime of the kernel, transposed to [in][out] so that consecutive threads (consecutive
// outputs) read consecutive words; activations ping-pong between two [rows][128] tiles.
// fp32 multiply-add (fmaf), fp32 accumulate, bias first then inputs ascending — the same
// summation order as the oracle (slimnn/src/linear.rs:17-25) up to FMA rounding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mlp {

constexpr int XS = 128; // activation row stride (floats)
// transposed weight offsets in the shared weight block (floats); layer 1 input padded 63 -> 64
constexpr int W1 = 0;                 // [64][128]
constexpr int W2 = W1 + 64 * 128;     // [128][96]
constexpr int W3 = W2 + 128 * 96;     // [96][64]
constexpr int W4 = W3 + 96 * 64;      // [64][48]
constexpr int W5 = W4 + 64 * 48;      // [48][12]
constexpr int B1 = W5 + 48 * 12;      // biases
constexpr int B2 = B1 + 128;
constexpr int B3 = B2 + 96;
constexpr int B4 = B3 + 64;
constexpr int B5 = B4 + 48;
constexpr int WEIGHT_FLOATS = ((B5 + 12 + 3) / 4) * 4; // 30620 -> multiple of 4

// blob order (synthesis_b200.h): l_k.weight [out][in], l_k.bias, k = 1..5
__device__ __forceinline__ void load_weights_transposed(float* sw, const float* __restrict__ blob, int tid, int nthreads) {
    const int I[5] = {63, 128, 96, 64, 48}, O[5] = {128, 96, 64, 48, 12};
    const int WO[5] = {W1, W2, W3, W4, W5}, BO[5] = {B1, B2, B3, B4, B5};
    for (int i = tid; i < WEIGHT_FLOATS; i += nthreads) sw[i] = 0.0f;
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        int n = I[l] * O[l];
        for (int e = tid; e < n; e += nthreads) {
            int o = e / I[l], i = e - o * I[l];
            sw[WO[l] + i * O[l] + o] = blob[off + e];
        }
        off += n;
        for (int e = tid; e < O[l]; e += nthreads) sw[BO[l] + e] = blob[off + e];
        off += O[l];
    }
    __syncthreads();
}

// One layer for R rows with T threads: thread -> (output o, row group); each thread walks the
// inputs four at a time (one float4 activation read per row, four weight reads shared by its rows).
template <int I, int O, int OP, bool RELU, int R, int T>
__device__ __forceinline__ void layer(const float* __restrict__ Wt, const float* __restrict__ bias, const float* xin, float* xout, int tid) {
    constexpr int NG = T / OP;
    constexpr int RPT = (R + NG - 1) / NG;
    const int o = tid % OP, rg = tid / OP;
    if (o < O && rg * RPT < R) {
        float acc[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) acc[r] = bias[o];
#pragma unroll 4
        for (int i = 0; i < I; i += 4) {
            float w0 = Wt[(i + 0) * O + o], w1 = Wt[(i + 1) * O + o], w2 = Wt[(i + 2) * O + o], w3 = Wt[(i + 3) * O + o];
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                int row = rg * RPT + r;
                if (row < R) {
                    float4 x = *reinterpret_cast<const float4*>(xin + row * XS + i);
                    acc[r] = fmaf(x.x, w0, acc[r]);
                    acc[r] = fmaf(x.y, w1, acc[r]);
                    acc[r] = fmaf(x.z, w2, acc[r]);
                    acc[r] = fmaf(x.w, w3, acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            int row = rg * RPT + r;
            if (row < R) xout[row * XS + o] = RELU ? fmaxf(acc[r], 0.0f) : acc[r];
        }
    }
}

// Whole forward: input in xa[row][0..63] (feature 63 = 0 padding), result in xb[row][0..11]
// (9 policy logits, 3 value logits).  All T threads of the CTA must call it.
template <int R, int T>
__device__ __forceinline__ void forward(const float* sw, float* xa, float* xb, int tid) {
    layer<64, 128, 128, true, R, T>(sw + W1, sw + B1, xa, xb, tid);
    __syncthreads();
    layer<128, 96, 128, true, R, T>(sw + W2, sw + B2, xb, xa, tid);
    __syncthreads();
    layer<96, 64, 64, true, R, T>(sw + W3, sw + B3, xa, xb, tid);
    __syncthreads();
    layer<64, 48, 64, true, R, T>(sw + W4, sw + B4, xb, xa, tid);
    __syncthreads();
    layer<48, 12, 16, false, R, T>(sw + W5, sw + B5, xa, xb, tid);
    __syncthreads();
}

} // namespace mlp
