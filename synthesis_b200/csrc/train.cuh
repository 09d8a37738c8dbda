// train.cuh — the learner's inner loop on the device: forward, KL losses, backward and Adam for Connect4Net,
// one optimizer step per batch of 32 rows, any number of steps per launch.
//
// Replaces synthesis/src/alpha_zero.rs:73-92 of the reference (per batch: policy.forward, log_softmax, kl_div with
// Reduction::Sum times 1/batch_size, weighted sum, opt.backward_step) with the network of
// study-connect4/src/policies.rs:13-46 and tch's Adam (libtorch torch::optim::Adam::step: bias corrections in double,
// denom = sqrt(v)/sqrt(bc2) + eps, p -= lr/bc1 * m/denom, L2 weight decay added to the gradient).
//
// SGD steps depend on each other, so the loop is latency-bound by construction: ONE persistent CTA of 1024 threads keeps
// the fp32 master weights (padded rows, 127 KB), all activations and both delta buffers in shared memory for the whole
// launch; only the Adam moments stream through L2 (256 KB per step) and the next batch is prefetched into registers
// while the current step runs.  Features are synthesised from the bitboards (16 bytes per row instead of 252).
// After the last step the weights go back to the engine's blob, so the next gather searches with them without a host
// round trip (the reference writes model_{i}.ot and every worker reloads it, alpha_zero.rs:102, 192-194).
#pragma once
#include <stdint.h>

#include "c4.cuh"

namespace trn {

constexpr int NT = 1024;
constexpr int B = 32;   // batch size (LearningConfig::batch_size of study-connect4/src/main.rs:21)
constexpr int DS = 128; // row stride of the delta buffers

struct Layer { int in, out, kp, ld, woff, boff, blob_w, blob_b; };
// kp = in rounded up to 4, ld = kp + 4 (rows 16-byte aligned, conflict-free for 128-bit loads across output neurons)
constexpr Layer LAYERS[5] = {
    {63, 128, 64, 68, 0, 0, 0, 8064},
    {128, 96, 128, 132, 8704, 128, 8192, 20480},
    {96, 64, 96, 100, 21376, 224, 20576, 26720},
    {64, 48, 64, 68, 27776, 288, 26784, 29856},
    {48, 12, 48, 52, 31040, 336, 29904, 30480},
};
constexpr int W_FLOATS = 31040 + 12 * 52; // 31664
constexpr int B_FLOATS = 352;             // 348 biases, padded
constexpr int P_FLOATS = W_FLOATS + B_FLOATS; // the padded parameter vector; Adam's m and v use the same indexing
constexpr int N_BLOB = 30492;
// activations: A0 = features [32][64], A1 [32][128], A2 [32][96], A3 [32][64], A4 [32][48], A5 = logits [32][16]
__host__ __device__ constexpr int act_stride(int l) { return l == 0 ? 64 : l == 1 ? 128 : l == 2 ? 96 : l == 3 ? 64 : l == 4 ? 48 : 16; }
__host__ __device__ constexpr int act_off(int l) { return l == 0 ? 0 : l == 1 ? 2048 : l == 2 ? 6144 : l == 3 ? 9216 : l == 4 ? 11264 : 12800; }
constexpr int A_FLOATS = 12800 + 32 * 16;

struct Smem {
    float w[W_FLOATS];
    float bias[B_FLOATS];
    float act[A_FLOATS];
    float delta[2][B * DS];
    float target[B][12];
    float loss[2];
};

struct Params {
    float* blob;               // [30492] in/out: l_1.weight .. l_5.bias, PyTorch [out][in]
    float* m;                  // [P_FLOATS] Adam exp_avg
    float* v;                  // [P_FLOATS] Adam exp_avg_sq
    const uint64_t* my;        // FlatBatch rows
    const uint64_t* op;
    const float* pis;          // [n][9]
    const float* vs;           // [n][3]
    const uint32_t* batch_idx; // [n_steps][32 * micro] row indices
    const float2* sched;       // [n_steps] {lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)}
    float* losses;             // [n_steps][2] {pi_loss, v_loss} or null
    int* error;
    unsigned long long* prof;  // optional [16]: cycles per phase of a step, summed over steps (rank 0, thread 0; SYN_TRAIN_PROF=1)
    uint32_t n_rows, n_steps;
    uint32_t micro;            // micro-batches of 32 rows per optimizer step (batch_size / 32): gradients are accumulated over them
    float beta1, beta2, eps, wd, pw, vw;
};

template <int L>
__device__ __forceinline__ void forward(Smem& s) {
    constexpr Layer ly = LAYERS[L];
    constexpr int K4 = ly.kp / 4, AS = act_stride(L), OS = act_stride(L + 1);
    const int t = threadIdx.x;
    if (t < ly.out * 8) {
        const int o = t % ly.out, bg = t / ly.out;
        const float4* w = reinterpret_cast<const float4*>(s.w + ly.woff + o * ly.ld);
        const float4* a = reinterpret_cast<const float4*>(s.act + act_off(L) + (bg * 4) * AS);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int k = 0; k < K4; ++k) {
            const float4 wv = w[k];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 av = a[j * (AS / 4) + k];
                acc[j] = fmaf(wv.x, av.x, acc[j]);
                acc[j] = fmaf(wv.y, av.y, acc[j]);
                acc[j] = fmaf(wv.z, av.z, acc[j]);
                acc[j] = fmaf(wv.w, av.w, acc[j]);
            }
        }
        const float bv = s.bias[ly.boff + o];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float z = acc[j] + bv;
            if (L < 4) z = fmaxf(z, 0.0f);
            s.act[act_off(L + 1) + (bg * 4 + j) * OS + o] = z;
        }
    }
}

// delta of layer L-1's output = (W_L^T delta_L) masked by relu'(A_L); din = delta_L, dout = delta_{L-1}
template <int L>
__device__ __forceinline__ void backward_input(Smem& s, const float* din, float* dout) {
    constexpr Layer ly = LAYERS[L];
    constexpr int I4 = ly.kp / 4, AS = act_stride(L);
    const int t = threadIdx.x;
    if (t < I4 * 16) {
        const int i4 = t % I4, b0 = (t / I4) * 2;
        float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
        const float* w = s.w + ly.woff + 4 * i4;
#pragma unroll 4
        for (int o = 0; o < ly.out; ++o) {
            const float4 wv = *reinterpret_cast<const float4*>(w + o * ly.ld);
            const float d0 = din[b0 * DS + o], d1 = din[(b0 + 1) * DS + o];
            a0[0] = fmaf(d0, wv.x, a0[0]); a0[1] = fmaf(d0, wv.y, a0[1]); a0[2] = fmaf(d0, wv.z, a0[2]); a0[3] = fmaf(d0, wv.w, a0[3]);
            a1[0] = fmaf(d1, wv.x, a1[0]); a1[1] = fmaf(d1, wv.y, a1[1]); a1[2] = fmaf(d1, wv.z, a1[2]); a1[3] = fmaf(d1, wv.w, a1[3]);
        }
        const float4 x0 = *reinterpret_cast<const float4*>(s.act + act_off(L) + b0 * AS + 4 * i4);
        const float4 x1 = *reinterpret_cast<const float4*>(s.act + act_off(L) + (b0 + 1) * AS + 4 * i4);
        *reinterpret_cast<float4*>(dout + b0 * DS + 4 * i4) =
            make_float4(x0.x > 0.f ? a0[0] : 0.f, x0.y > 0.f ? a0[1] : 0.f, x0.z > 0.f ? a0[2] : 0.f, x0.w > 0.f ? a0[3] : 0.f);
        *reinterpret_cast<float4*>(dout + (b0 + 1) * DS + 4 * i4) =
            make_float4(x1.x > 0.f ? a1[0] : 0.f, x1.y > 0.f ? a1[1] : 0.f, x1.z > 0.f ? a1[2] : 0.f, x1.w > 0.f ? a1[3] : 0.f);
    }
}

struct Grad { float g[2][8]; float gb; };

// dW tiles of 8 (columns 4c..4c+3 and kp/2+4c..+3 of one output row) and the bias gradient of row threadIdx.x
template <int L>
__device__ __forceinline__ void weight_grad(const Smem& s, const float* d, Grad& gr) {
    constexpr Layer ly = LAYERS[L];
    constexpr int C8 = ly.kp / 8, TILES = ly.out * C8, AS = act_stride(L);
    const float* a = s.act + act_off(L);
#pragma unroll
    for (int slot = 0; slot < 2; ++slot) {
        const int tile = threadIdx.x + slot * NT;
        if (slot * NT < TILES && tile < TILES) {
            const int o = tile / C8, c = tile % C8;
            float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
            for (int b = 0; b < B; ++b) {
                const float dv = d[b * DS + o];
                const float4 x0 = *reinterpret_cast<const float4*>(a + b * AS + 4 * c);
                const float4 x1 = *reinterpret_cast<const float4*>(a + b * AS + ly.kp / 2 + 4 * c);
                g[0] = fmaf(dv, x0.x, g[0]); g[1] = fmaf(dv, x0.y, g[1]); g[2] = fmaf(dv, x0.z, g[2]); g[3] = fmaf(dv, x0.w, g[3]);
                g[4] = fmaf(dv, x1.x, g[4]); g[5] = fmaf(dv, x1.y, g[5]); g[6] = fmaf(dv, x1.z, g[6]); g[7] = fmaf(dv, x1.w, g[7]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) gr.g[slot][k] = g[k];
        }
    }
    if (threadIdx.x < ly.out) {
        float sum = 0.f;
#pragma unroll 8
        for (int b = 0; b < B; ++b) sum += d[b * DS + threadIdx.x];
        gr.gb = sum;
    }
}

// torch::optim::Adam::step on one parameter
__device__ __forceinline__ float adam1(const Params& p, float2 sc, float w, float g, float& m, float& v) {
    if (p.wd != 0.0f) g = fmaf(p.wd, w, g);
    m = fmaf(p.beta1, m, (1.0f - p.beta1) * g);
    v = fmaf(p.beta2, v, (1.0f - p.beta2) * g * g);
    // sc.y = 1 / sqrt(bias_correction2): one IEEE square root, one correctly rounded reciprocal, no division sequences
    const float denom = __fsqrt_rn(v) * sc.y + p.eps;
    return w - sc.x * (m * __frcp_rn(denom));
}

__device__ __forceinline__ void adam4(const Params& p, float2 sc, float* w, int idx, const float* g) {
    float4 wv = *reinterpret_cast<float4*>(w + idx);
    float4 mv = *reinterpret_cast<const float4*>(p.m + idx), vv = *reinterpret_cast<const float4*>(p.v + idx);
    wv.x = adam1(p, sc, wv.x, g[0], mv.x, vv.x);
    wv.y = adam1(p, sc, wv.y, g[1], mv.y, vv.y);
    wv.z = adam1(p, sc, wv.z, g[2], mv.z, vv.z);
    wv.w = adam1(p, sc, wv.w, g[3], mv.w, vv.w);
    *reinterpret_cast<float4*>(w + idx) = wv;
    *reinterpret_cast<float4*>(p.m + idx) = mv;
    *reinterpret_cast<float4*>(p.v + idx) = vv;
}

template <int L>
__device__ __forceinline__ void adam_layer(const Params& p, Smem& s, float2 sc, const Grad& gr) {
    constexpr Layer ly = LAYERS[L];
    constexpr int C8 = ly.kp / 8, TILES = ly.out * C8;
#pragma unroll
    for (int slot = 0; slot < 2; ++slot) {
        const int tile = threadIdx.x + slot * NT;
        if (slot * NT < TILES && tile < TILES) {
            const int o = tile / C8, c = tile % C8;
            const int idx = ly.woff + o * ly.ld + 4 * c;
            adam4(p, sc, s.w, idx, gr.g[slot]);
            adam4(p, sc, s.w, idx + ly.kp / 2, gr.g[slot] + 4);
        }
    }
    if (threadIdx.x < ly.out) {
        const int bi = ly.boff + threadIdx.x;
        float m = p.m[W_FLOATS + bi], v = p.v[W_FLOATS + bi];
        s.bias[bi] = adam1(p, sc, s.bias[bi], gr.gb, m, v);
        p.m[W_FLOATS + bi] = m;
        p.v[W_FLOATS + bi] = v;
    }
}

// one backward phase: delta_{L-1} (if L > 0), gradients of layer L; barrier; Adam on layer L
template <int L>
__device__ __forceinline__ void backward_layer(const Params& p, Smem& s, float2 sc, const float* din, float* dout) {
    if (L > 0) backward_input<L>(s, din, dout);
    Grad gr;
    weight_grad<L>(s, din, gr);
    __syncthreads();
    adam_layer<L>(p, s, sc, gr);
}

// blob (PyTorch [out][in]) <-> padded rows in shared memory
template <int L, bool LOAD>
__device__ __forceinline__ void blob_layer(const Params& p, Smem& s) {
    constexpr Layer ly = LAYERS[L];
    const int t = threadIdx.x;
    for (int i = t; i < ly.out * ly.in; i += NT) {
        const int si = ly.woff + (i / ly.in) * ly.ld + (i % ly.in);
        if (LOAD) s.w[si] = p.blob[ly.blob_w + i];
        else p.blob[ly.blob_w + i] = s.w[si];
    }
    if (t < ly.out) {
        if (LOAD) s.bias[ly.boff + t] = p.blob[ly.blob_b + t];
        else p.blob[ly.blob_b + t] = s.bias[ly.boff + t];
    }
}

__global__ void __launch_bounds__(NT, 1) train_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, b = t >> 5;
    // ---- weights: blob -> padded rows (pads zero, they stay zero: their gradients are zero)
    for (int i = t; i < W_FLOATS + B_FLOATS; i += NT) s.w[i] = 0.0f; // w and bias are adjacent
    __syncthreads();
    blob_layer<0, true>(p, s); blob_layer<1, true>(p, s); blob_layer<2, true>(p, s); blob_layer<3, true>(p, s); blob_layer<4, true>(p, s);
    // ---- first batch into registers
    uint64_t pf_my = 0, pf_op = 0;
    float pf_t = 0.0f;
    auto prefetch = [&](uint32_t step) {
        uint32_t idx = p.batch_idx[(size_t)step * B + b];
        if (idx >= p.n_rows) { atomicExch(p.error, 1); idx = 0; }
        pf_my = p.my[idx];
        pf_op = p.op[idx];
        pf_t = lane < 9 ? p.pis[(size_t)idx * 9 + lane] : (lane < 12 ? p.vs[(size_t)idx * 3 + (lane - 9)] : 0.0f);
    };
    if (p.n_steps) prefetch(0);
    __syncthreads();
    for (uint32_t step = 0; step < p.n_steps; ++step) {
        // ---- batch: features (Game::features, connect4.rs:237-258) and targets
        s.act[act_off(0) + b * 64 + lane] = c4::feature(pf_my, pf_op, lane);
        s.act[act_off(0) + b * 64 + 32 + lane] = lane < 31 ? c4::feature(pf_my, pf_op, 32 + lane) : 0.0f;
        if (lane < 12) s.target[b][lane] = pf_t;
        const float2 sc = p.sched[step];
        if (step + 1u < p.n_steps) prefetch(step + 1u);
        __syncthreads();
        forward<0>(s); __syncthreads();
        forward<1>(s); __syncthreads();
        forward<2>(s); __syncthreads();
        forward<3>(s); __syncthreads();
        forward<4>(s); __syncthreads();
        // ---- log_softmax + kl_div(Reduction::Sum) / batch for both heads; delta of the logits
        if (t < 64) {
            const int head = t >> 5, n = head ? 3 : 9, o0 = head ? 9 : 0;
            const float* z = s.act + act_off(5) + lane * 16 + o0;
            const float* tg = &s.target[lane][o0];
            float mx = z[0];
            for (int k = 1; k < n; ++k) mx = fmaxf(mx, z[k]);
            float se = 0.f, st = 0.f;
            for (int k = 0; k < n; ++k) { se += expf(z[k] - mx); st += tg[k]; }
            const float lse = mx + logf(se);
            const float scale = (head ? p.vw : p.pw) * (1.0f / (float)B);
            float loss = 0.f;
            for (int k = 0; k < n; ++k) {
                const float lp = z[k] - lse, tk = tg[k];
                if (tk > 0.f) loss += tk * (logf(tk) - lp); // kl_div: xlogy(t, t) - t * input
                s.delta[0][lane * DS + o0 + k] = scale * (expf(lp) * st - tk);
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, d);
            if (lane == 0 && p.losses) p.losses[(size_t)step * 2 + head] = loss * (1.0f / (float)B);
        }
        __syncthreads();
        backward_layer<4>(p, s, sc, s.delta[0], s.delta[1]);
        backward_layer<3>(p, s, sc, s.delta[1], s.delta[0]);
        backward_layer<2>(p, s, sc, s.delta[0], s.delta[1]);
        backward_layer<1>(p, s, sc, s.delta[1], s.delta[0]);
        backward_layer<0>(p, s, sc, s.delta[0], s.delta[1]);
        __syncthreads();
    }
    // ---- weights back to the blob
    blob_layer<0, false>(p, s); blob_layer<1, false>(p, s); blob_layer<2, false>(p, s); blob_layer<3, false>(p, s); blob_layer<4, false>(p, s);
}

} // namespace trn
