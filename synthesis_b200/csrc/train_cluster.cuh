// train_cluster.cuh — the learner's batch loop (train.cuh) spread over a thread-block CLUSTER of 8 SMs.
//
// Same algorithm and interface as trn::train_kernel (alpha_zero.rs:73-92, torch::optim::Adam::step); what changes is
// who owns what.  CTA `rank` of the cluster owns a slice of every layer's OUTPUT neurons (16, 12, 8, 6, 2 rows): their
// fp32 master weights AND their Adam moments live in its shared memory for the whole launch (48 KB — nothing streams
// through L2 any more), it computes their activations, their weight gradients and their updates.  Two exchanges per
// layer go through distributed shared memory:
//   forward : a CTA stores its slice of the layer's activations into all 8 CTAs' activation buffers (all-gather by
//             remote stores), then one cluster barrier;
//   backward: delta_{l-1} = W_l^T delta_l sums over ALL output rows, so every CTA writes the partial sum over ITS rows to
//             a local buffer, one cluster barrier, and each CTA adds the 8 partials of its own slice with remote loads
//             (reduce-scatter in fixed rank order, so results do not depend on timing).
// A step is 10 cluster barriers and about a tenth of the single-CTA kernel's dependent instruction stream.
#pragma once
#include <cooperative_groups.h>
#include <stdint.h>

#include "train.cuh"

namespace trc {

namespace cg = cooperative_groups;
using trn::act_off;
using trn::act_stride;
using trn::B;
using trn::Layer;
using trn::LAYERS;

constexpr int NC = 8;   // CTAs per cluster (the portable maximum)
constexpr int NT = 512; // threads per CTA
constexpr int DSL = 16; // row stride of the own-delta buffer (the widest slice)
constexpr int PS = 128; // row stride of the partial buffers

__host__ __device__ constexpr int slice(int l) { return (LAYERS[l].out + NC - 1) / NC; } // 16, 12, 8, 6, 2
constexpr int OWN_W = 16 * 68 + 12 * 132 + 8 * 100 + 6 * 68 + 2 * 52; // 3984
constexpr int OWN_P = OWN_W + 48;                                    // + 44 biases, padded

struct Smem {
    float w[OWN_P], m[OWN_P], v[OWN_P], g[OWN_P]; // g: gradient sums of a batch larger than 32 rows (async kernel only)
    float act[trn::A_FLOATS]; // ALL activations of the batch (filled by every CTA's remote stores)
    float part[2][B * PS];    // partial input-deltas over the own rows, double-buffered by layer parity
    float d[B * DSL];         // delta of the own output slice of the current layer
    float target[B][12];
};

__device__ __forceinline__ int rows_owned(int l_out, int sl, unsigned rank) {
    const int lo = (int)rank * sl;
    return max(0, min(sl, l_out - lo));
}

// own rows of layer L: where they live in Smem::w / m / v
template <int L>
struct Own {
    static constexpr int sl = (LAYERS[L].out + NC - 1) / NC;
    static constexpr int woff = L == 0 ? 0 : L == 1 ? 1088 : L == 2 ? 2672 : L == 3 ? 3472 : 3880;
    static constexpr int boff = OWN_W + (L == 0 ? 0 : L == 1 ? 16 : L == 2 ? 28 : L == 3 ? 36 : 42);
};

// blob + Adam moments in global memory (train.cuh's padded layout) <-> the own rows
template <int L, bool LOAD, class S>
__device__ __forceinline__ void sync_rows(const trn::Params& p, S& s, unsigned rank) {
    constexpr Layer ly = LAYERS[L];
    constexpr int sl = Own<L>::sl;
    const int lo = (int)rank * sl, n = rows_owned(ly.out, sl, rank);
    for (int i = threadIdx.x; i < n * ly.kp; i += NT) {
        const int r = i / ly.kp, c = i % ly.kp;
        const int si = Own<L>::woff + r * ly.ld + c, gi = ly.woff + (lo + r) * ly.ld + c;
        if (LOAD) {
            s.w[si] = c < ly.in ? p.blob[ly.blob_w + (lo + r) * ly.in + c] : 0.0f;
            s.m[si] = p.m[gi];
            s.v[si] = p.v[gi];
        } else {
            if (c < ly.in) p.blob[ly.blob_w + (lo + r) * ly.in + c] = s.w[si];
            p.m[gi] = s.m[si];
            p.v[gi] = s.v[si];
        }
    }
    if ((int)threadIdx.x < n) {
        const int si = Own<L>::boff + threadIdx.x, gi = trn::W_FLOATS + ly.boff + lo + threadIdx.x;
        if (LOAD) { s.w[si] = p.blob[ly.blob_b + lo + threadIdx.x]; s.m[si] = p.m[gi]; s.v[si] = p.v[gi]; }
        else { p.blob[ly.blob_b + lo + threadIdx.x] = s.w[si]; p.m[gi] = s.m[si]; p.v[gi] = s.v[si]; }
    }
}

// own slice of layer L's outputs for all 32 rows, stored into every CTA's activation buffer
template <int L>
__device__ __forceinline__ void forward(cg::cluster_group& cluster, Smem& s, unsigned rank) {
    constexpr Layer ly = LAYERS[L];
    constexpr int sl = Own<L>::sl, K4 = ly.kp / 4, AS = act_stride(L), OS = act_stride(L + 1);
    const int t = threadIdx.x;
    const int ol = t % sl, b = t / sl, og = (int)rank * sl + ol;
    if (t < sl * B && og < ly.out) {
        const float4* w = reinterpret_cast<const float4*>(s.w + Own<L>::woff + ol * ly.ld);
        const float4* a = reinterpret_cast<const float4*>(s.act + act_off(L) + b * AS);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
        for (int k = 0; k < K4; ++k) {
            const float4 wv = w[k], av = a[k];
            a0 = fmaf(wv.x, av.x, a0); a1 = fmaf(wv.y, av.y, a1); a2 = fmaf(wv.z, av.z, a2); a3 = fmaf(wv.w, av.w, a3);
        }
        float z = ((a0 + a1) + (a2 + a3)) + s.w[Own<L>::boff + ol];
        if (L < 4) z = fmaxf(z, 0.0f);
        const int at = act_off(L + 1) + b * OS + og;
#pragma unroll
        for (unsigned r = 0; r < (unsigned)NC; ++r) cluster.map_shared_rank(s.act, r)[at] = z;
    }
}

// partial[b][i] = sum over the own rows o of delta[b][o] * W[o][i]
template <int L>
__device__ __forceinline__ void backward_partial(Smem& s, unsigned rank) {
    constexpr Layer ly = LAYERS[L];
    constexpr int sl = Own<L>::sl, I4 = ly.kp / 4, TILES = B * I4;
    const int n = rows_owned(ly.out, sl, rank);
    float* part = s.part[L & 1];
    for (int tile = threadIdx.x; tile < TILES; tile += NT) {
        const int i4 = tile % I4, b = tile / I4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int o = 0; o < n; ++o) {
            const float4 wv = *reinterpret_cast<const float4*>(s.w + Own<L>::woff + o * ly.ld + 4 * i4);
            const float dv = s.d[b * DSL + o];
            acc.x = fmaf(dv, wv.x, acc.x); acc.y = fmaf(dv, wv.y, acc.y); acc.z = fmaf(dv, wv.z, acc.z); acc.w = fmaf(dv, wv.w, acc.w);
        }
        *reinterpret_cast<float4*>(part + b * PS + 4 * i4) = acc;
    }
}

// gradient of four weights of an own row (and of the row's bias), then Adam on them.
// accum: 0 = a batch of 32 rows (the gradient is complete); 1 = a micro-batch of a larger batch: add to s.g, no update;
// 2 = its last micro-batch: add, update with the sum, clear s.g.  (loss = kl_div(sum) / batch_size, alpha_zero.rs:80-88: the
// per-row deltas already carry 1 / batch_size, so the sum over micro-batches IS the batch gradient.)
template <int L, class S>
__device__ __forceinline__ void grad_and_adam(const trn::Params& p, S& s, float2 sc, unsigned rank, int accum = 0) {
    constexpr Layer ly = LAYERS[L];
    constexpr int sl = Own<L>::sl, C4 = ly.kp / 4, AS = act_stride(L);
    static_assert(sl * C4 <= NT, "one tile per thread");
    const int n = rows_owned(ly.out, sl, rank);
    const int t = threadIdx.x, ol = t / C4, c = t % C4;
    const bool tile = t < n * C4, brow = t < n;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    float gb = 0.f;
    if (tile) {
        const float* a = s.act + act_off(L) + 4 * c;
#pragma unroll 8
        for (int b = 0; b < B; ++b) {
            const float dv = s.d[b * DSL + ol];
            const float4 x = *reinterpret_cast<const float4*>(a + b * AS);
            g.x = fmaf(dv, x.x, g.x); g.y = fmaf(dv, x.y, g.y); g.z = fmaf(dv, x.z, g.z); g.w = fmaf(dv, x.w, g.w);
        }
    }
    if (brow) {
#pragma unroll 8
        for (int b = 0; b < B; ++b) gb += s.d[b * DSL + t];
    }
    __syncthreads(); // backward_partial<L> has read the weights this is about to change
    if (accum != 0) { // the own tile's running sums live at the tile's own index: no other thread touches them
        if (tile) {
            const int idx = Own<L>::woff + ol * ly.ld + 4 * c;
            const float4 pg = *reinterpret_cast<const float4*>(s.g + idx);
            g.x += pg.x; g.y += pg.y; g.z += pg.z; g.w += pg.w;
            *reinterpret_cast<float4*>(s.g + idx) = accum == 2 ? make_float4(0.f, 0.f, 0.f, 0.f) : g;
        }
        if (brow) {
            const int bi = Own<L>::boff + t;
            gb += s.g[bi];
            s.g[bi] = accum == 2 ? 0.f : gb;
        }
        if (accum == 1) return;
    }
    if (tile) {
        const int idx = Own<L>::woff + ol * ly.ld + 4 * c;
        float4 wv = *reinterpret_cast<float4*>(s.w + idx), mv = *reinterpret_cast<float4*>(s.m + idx), vv = *reinterpret_cast<float4*>(s.v + idx);
        wv.x = trn::adam1(p, sc, wv.x, g.x, mv.x, vv.x);
        wv.y = trn::adam1(p, sc, wv.y, g.y, mv.y, vv.y);
        wv.z = trn::adam1(p, sc, wv.z, g.z, mv.z, vv.z);
        wv.w = trn::adam1(p, sc, wv.w, g.w, mv.w, vv.w);
        *reinterpret_cast<float4*>(s.w + idx) = wv;
        *reinterpret_cast<float4*>(s.m + idx) = mv;
        *reinterpret_cast<float4*>(s.v + idx) = vv;
    }
    if (brow) {
        const int bi = Own<L>::boff + t;
        s.w[bi] = trn::adam1(p, sc, s.w[bi], gb, s.m[bi], s.v[bi]);
    }
}

// own slice of delta_{L-1}: the 8 partials of layer L in rank order, masked by relu'(A_L)
template <int L>
__device__ __forceinline__ void reduce_scatter(cg::cluster_group& cluster, Smem& s, unsigned rank) {
    constexpr Layer lp = LAYERS[L - 1];
    constexpr int sl = Own<L - 1>::sl, AS = act_stride(L);
    const int t = threadIdx.x, j = t % sl, b = t / sl, ig = (int)rank * sl + j;
    if (t < sl * B && ig < lp.out) {
        const int at = (L & 1) * (B * PS) + b * PS + ig;
        float sum = 0.f;
        float x[NC];
#pragma unroll
        for (unsigned r = 0; r < (unsigned)NC; ++r) x[r] = cluster.map_shared_rank(&s.part[0][0], r)[at];
#pragma unroll
        for (int r = 0; r < NC; ++r) sum += x[r];
        s.d[b * DSL + j] = s.act[act_off(L) + b * AS + ig] > 0.f ? sum : 0.f;
    }
}

template <int L>
__device__ __forceinline__ void backward_layer(cg::cluster_group& cluster, const trn::Params& p, Smem& s, float2 sc, unsigned rank) {
    if (L > 0) backward_partial<L>(s, rank);
    grad_and_adam<L>(p, s, sc, rank);
    if (L > 0) {
        cluster.sync(); // every CTA's partial of layer L is complete (and s.d has been consumed everywhere)
        reduce_scatter<(L > 0 ? L : 1)>(cluster, s, rank);
        __syncthreads();
    }
}

__global__ void __cluster_dims__(NC, 1, 1) __launch_bounds__(NT, 1) train_cluster_kernel(const __grid_constant__ trn::Params p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int t = threadIdx.x, lane = t & 31;
    for (int i = t; i < 4 * OWN_P; i += NT) s.w[i] = 0.0f; // w, m, v, g are adjacent
    __syncthreads();
    sync_rows<0, true>(p, s, rank); sync_rows<1, true>(p, s, rank); sync_rows<2, true>(p, s, rank);
    sync_rows<3, true>(p, s, rank); sync_rows<4, true>(p, s, rank);
    // the batch: warp w of 16 loads rows w and w + 16 (every CTA loads the whole batch — 2 KB — for itself)
    uint64_t pf_my[2] = {0, 0}, pf_op[2] = {0, 0};
    float pf_t[2] = {0.f, 0.f};
    auto prefetch = [&](uint32_t step) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t idx = p.batch_idx[(size_t)step * B + (t >> 5) + 16 * h];
            if (idx >= p.n_rows) { atomicExch(p.error, 1); idx = 0; }
            pf_my[h] = p.my[idx];
            pf_op[h] = p.op[idx];
            pf_t[h] = lane < 9 ? p.pis[(size_t)idx * 9 + lane] : (lane < 12 ? p.vs[(size_t)idx * 3 + (lane - 9)] : 0.0f);
        }
    };
    if (p.n_steps) prefetch(0);
    cluster.sync(); // every CTA is resident and initialised before the first remote store
    for (uint32_t step = 0; step < p.n_steps; ++step) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b = (t >> 5) + 16 * h;
            s.act[act_off(0) + b * 64 + lane] = c4::feature(pf_my[h], pf_op[h], lane);
            s.act[act_off(0) + b * 64 + 32 + lane] = lane < 31 ? c4::feature(pf_my[h], pf_op[h], 32 + lane) : 0.0f;
            if (lane < 12) s.target[b][lane] = pf_t[h];
        }
        const float2 sc = p.sched[step];
        if (step + 1u < p.n_steps) prefetch(step + 1u);
        __syncthreads();
        forward<0>(cluster, s, rank); cluster.sync();
        forward<1>(cluster, s, rank); cluster.sync();
        forward<2>(cluster, s, rank); cluster.sync();
        forward<3>(cluster, s, rank); cluster.sync();
        forward<4>(cluster, s, rank); cluster.sync();
        // losses and the delta of the logits: every CTA for the whole batch (64 threads), keeping its own two columns
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
            const int lo = (int)rank * Own<4>::sl;
            float loss = 0.f;
            for (int k = 0; k < n; ++k) {
                const float lp = z[k] - lse, tk = tg[k];
                if (tk > 0.f) loss += tk * (logf(tk) - lp);
                const int o = o0 + k;
                if (o >= lo && o < lo + Own<4>::sl) s.d[lane * DSL + (o - lo)] = scale * (expf(lp) * st - tk);
            }
#pragma unroll
            for (int dlt = 16; dlt; dlt >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, dlt);
            if (lane == 0 && rank == 0 && p.losses) p.losses[(size_t)step * 2 + head] = loss * (1.0f / (float)B);
        }
        __syncthreads();
        backward_layer<4>(cluster, p, s, sc, rank);
        backward_layer<3>(cluster, p, s, sc, rank);
        backward_layer<2>(cluster, p, s, sc, rank);
        backward_layer<1>(cluster, p, s, sc, rank);
        backward_layer<0>(cluster, p, s, sc, rank);
        cluster.sync(); // nobody still reads this step's activations when the next step's remote stores arrive
    }
    sync_rows<0, false>(p, s, rank); sync_rows<1, false>(p, s, rank); sync_rows<2, false>(p, s, rank);
    sync_rows<3, false>(p, s, rank); sync_rows<4, false>(p, s, rank);
}

// =====================================================================================================================
// The same schedule with ASYNCHRONOUS remote stores.  cluster.sync() costs a GPU-scope fence per exchange (SASS:
// MEMBAR.ALL.GPU + CCTL.IVALL around the hardware barrier); here every exchange is st.async into the consumer's shared
// memory, completing a transaction count on the CONSUMER's mbarrier (STAS.128), and the consumer waits on its own
// mbarrier — no fence, no cluster-wide rendezvous.  The input-delta exchange becomes a push too: a CTA sends each owner
// the partial sums for the owner's columns, the owner adds the 8 slices in rank order.  One relaxed hardware cluster
// barrier per step remains, only to keep the next step's stores from overtaking this step's readers.
struct SmemA {
    float w[OWN_P], m[OWN_P], v[OWN_P], g[OWN_P]; // g: gradient sums over the micro-batches of a batch larger than 32 rows
    float act[trn::A_FLOATS];
    float recv[2][NC][B * DSL]; // [layer parity][source rank]: partial input-deltas for the own columns
    float d[B * DSL];
    float target[B][12];
    float loss_row[B][2];
    uint64_t bar_act[5];  // activations of layer l + 1 complete (all 8 slices)
    uint64_t bar_part[4]; // partials of layer L from all 8 CTAs, index L - 1
};

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async4(uint32_t raddr, uint32_t rbar, float a, float b, float c, float d) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "f"(a), "f"(b), "f"(c),
                 "f"(d), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async2(uint32_t raddr, uint32_t rbar, float a, float b) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(raddr), "f"(a), "f"(b), "r"(rbar) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); }
__device__ __forceinline__ void bar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cluster_barrier_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}

// own slice of layer L's outputs; VW neighbouring lanes hand their value to a leader, which sends one vector per CTA
template <int L>
__device__ __forceinline__ void forward_async(SmemA& s, unsigned rank) {
    constexpr Layer ly = LAYERS[L];
    constexpr int sl = Own<L>::sl, K4 = ly.kp / 4, AS = act_stride(L), OS = act_stride(L + 1), VW = sl % 4 == 0 ? 4 : 2;
    const int t = threadIdx.x;
    const int ol = t % sl, b = t / sl, og = (int)rank * sl + ol;
    const bool valid = t < sl * B && og < ly.out;
    float z = 0.f;
    if (valid) {
        const float4* w = reinterpret_cast<const float4*>(s.w + Own<L>::woff + ol * ly.ld);
        const float4* a = reinterpret_cast<const float4*>(s.act + act_off(L) + b * AS);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
        for (int k = 0; k < K4; ++k) {
            const float4 wv = w[k], av = a[k];
            a0 = fmaf(wv.x, av.x, a0); a1 = fmaf(wv.y, av.y, a1); a2 = fmaf(wv.z, av.z, a2); a3 = fmaf(wv.w, av.w, a3);
        }
        z = ((a0 + a1) + (a2 + a3)) + s.w[Own<L>::boff + ol];
        if (L < 4) z = fmaxf(z, 0.0f);
    }
    const float z1 = __shfl_down_sync(0xffffffffu, z, 1), z2 = __shfl_down_sync(0xffffffffu, z, 2), z3 = __shfl_down_sync(0xffffffffu, z, 3);
    if (valid && (ol % VW) == 0) { // groups of VW never straddle a warp or a row (sl and 32 are multiples of VW)
        const uint32_t la = smem_u32(s.act + act_off(L + 1) + b * OS + og), lb = smem_u32(&s.bar_act[L]);
#pragma unroll
        for (uint32_t r = 0; r < (uint32_t)NC; ++r) {
            if (VW == 4) st_async4(mapa(la, r), mapa(lb, r), z, z1, z2, z3);
            else st_async2(mapa(la, r), mapa(lb, r), z, z1);
        }
    }
}

// partial input-deltas over the own rows, sent to the owner of each column group
template <int L>
__device__ __forceinline__ void push_partial(SmemA& s, unsigned rank) {
    constexpr Layer ly = LAYERS[L];
    constexpr int sl = Own<L>::sl, dsl = Own<L - 1>::sl, VW = dsl % 4 == 0 ? 4 : 2, IV = ly.in / VW, TILES = B * IV;
    const int n = rows_owned(ly.out, sl, rank);
    const uint32_t lrecv = smem_u32(&s.recv[L & 1][rank][0]), lbar = smem_u32(&s.bar_part[L - 1]);
    for (int tile = threadIdx.x; tile < TILES; tile += NT) {
        const int iv = tile % IV, b = tile / IV, i = iv * VW;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < n; ++o) {
            const float dv = s.d[b * DSL + o];
            const float* wp = s.w + Own<L>::woff + o * ly.ld + i;
            if (VW == 4) {
                const float4 wv = *reinterpret_cast<const float4*>(wp);
                acc[0] = fmaf(dv, wv.x, acc[0]); acc[1] = fmaf(dv, wv.y, acc[1]); acc[2] = fmaf(dv, wv.z, acc[2]); acc[3] = fmaf(dv, wv.w, acc[3]);
            } else {
                const float2 wv = *reinterpret_cast<const float2*>(wp);
                acc[0] = fmaf(dv, wv.x, acc[0]); acc[1] = fmaf(dv, wv.y, acc[1]);
            }
        }
        const uint32_t dest = (uint32_t)(i / dsl), j = (uint32_t)(i % dsl);
        const uint32_t ra = mapa(lrecv + (uint32_t)(b * DSL + j) * 4u, dest), rb = mapa(lbar, dest);
        if (VW == 4) st_async4(ra, rb, acc[0], acc[1], acc[2], acc[3]);
        else st_async2(ra, rb, acc[0], acc[1]);
    }
}

template <int L>
__device__ __forceinline__ void reduce_local(SmemA& s, unsigned rank) {
    constexpr int sl = Own<L - 1>::sl, AS = act_stride(L);
    const int t = threadIdx.x, j = t % sl, b = t / sl, ig = (int)rank * sl + j;
    if (t < sl * B) {
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < NC; ++r) sum += s.recv[L & 1][r][b * DSL + j];
        s.d[b * DSL + j] = s.act[act_off(L) + b * AS + ig] > 0.f ? sum : 0.f;
    }
}

// grad_and_adam works on any struct with w, m, v, d, act members
template <int L>
__device__ __forceinline__ void backward_layer_async(const trn::Params& p, SmemA& s, float2 sc, unsigned rank, uint32_t par, int accum) {
    if (L > 0) push_partial<(L > 0 ? L : 1)>(s, rank);
    grad_and_adam<L>(p, s, sc, rank, accum);
    if (L > 0) {
        bar_wait(&s.bar_part[(L > 0 ? L : 1) - 1], par);
        reduce_local<(L > 0 ? L : 1)>(s, rank);
        __syncthreads();
    }
}

__global__ void __cluster_dims__(NC, 1, 1) __launch_bounds__(NT, 1) train_cluster_async_kernel(const __grid_constant__ trn::Params p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SmemA& s = *reinterpret_cast<SmemA*>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int t = threadIdx.x, lane = t & 31;
    for (int i = t; i < 4 * OWN_P; i += NT) s.w[i] = 0.0f;
    if (t == 0) {
        for (int k = 0; k < 5; ++k) bar_init(&s.bar_act[k]);
        for (int k = 0; k < 4; ++k) bar_init(&s.bar_part[k]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    sync_rows<0, true>(p, s, rank); sync_rows<1, true>(p, s, rank); sync_rows<2, true>(p, s, rank);
    sync_rows<3, true>(p, s, rank); sync_rows<4, true>(p, s, rank);
    uint64_t pf_my[2] = {0, 0}, pf_op[2] = {0, 0};
    float pf_t[2] = {0.f, 0.f};
    uint32_t pf_idx[2] = {0u, 0u};
    const uint32_t n_micro = p.n_steps * p.micro; // micro-batches of 32 rows; batch_idx is [n_steps][32 * micro] = [n_micro][32]
    auto fetch_idx = [&](uint32_t ms) { // row indices first: the rows themselves are requested a whole micro-step later
#pragma unroll
        for (int h = 0; h < 2; ++h) pf_idx[h] = ms < n_micro ? p.batch_idx[(size_t)ms * B + (t >> 5) + 16 * h] : 0u;
    };
    auto prefetch = [&]() {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t idx = pf_idx[h];
            if (idx >= p.n_rows) { atomicExch(p.error, 1); idx = 0; }
            pf_my[h] = p.my[idx];
            pf_op[h] = p.op[idx];
            pf_t[h] = lane < 9 ? p.pis[(size_t)idx * 9 + lane] : (lane < 12 ? p.vs[(size_t)idx * 3 + (lane - 9)] : 0.0f);
        }
    };
    fetch_idx(0);
    if (p.n_steps) prefetch();
    fetch_idx(1);
    cluster.sync(); // every CTA resident, its mbarriers initialised and visible
    const bool prof = p.prof != nullptr && t == 0 && rank == 0;
    long long pc[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tk = prof ? clock64() : 0;
#define SYN_TICK(k) if (prof) { const long long now = clock64(); pc[k] += now - tk; tk = now; }
    const float inv_batch = 1.0f / (float)(B * p.micro);
    for (uint32_t ms = 0, step = 0, k_micro = 0; ms < n_micro; ++ms) {
        const uint32_t par = ms & 1u;
        const int accum = p.micro == 1u ? 0 : (k_micro + 1u == p.micro ? 2 : 1);
        if (t == 0) { // this step's phases: bytes that will land in THIS CTA
            bar_expect(&s.bar_act[0], 128u * 128u); bar_expect(&s.bar_act[1], 128u * 96u); bar_expect(&s.bar_act[2], 128u * 64u);
            bar_expect(&s.bar_act[3], 128u * 48u); bar_expect(&s.bar_act[4], 128u * 12u);
            bar_expect(&s.bar_part[3], 1024u * Own<3>::sl); bar_expect(&s.bar_part[2], 1024u * Own<2>::sl);
            bar_expect(&s.bar_part[1], 1024u * Own<1>::sl); bar_expect(&s.bar_part[0], 1024u * Own<0>::sl);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b = (t >> 5) + 16 * h;
            s.act[act_off(0) + b * 64 + lane] = c4::feature(pf_my[h], pf_op[h], lane);
            s.act[act_off(0) + b * 64 + 32 + lane] = lane < 31 ? c4::feature(pf_my[h], pf_op[h], 32 + lane) : 0.0f;
            if (lane < 12) s.target[b][lane] = pf_t[h];
        }
        const float2 sc = p.sched[step];
        if (ms + 1u < n_micro) prefetch(); // rows of the next micro-step (indices fetched during the previous one)
        fetch_idx(ms + 2u);
        __syncthreads();
        SYN_TICK(0)
        forward_async<0>(s, rank); bar_wait(&s.bar_act[0], par); SYN_TICK(1)
        forward_async<1>(s, rank); bar_wait(&s.bar_act[1], par); SYN_TICK(2)
        forward_async<2>(s, rank); bar_wait(&s.bar_act[2], par); SYN_TICK(3)
        forward_async<3>(s, rank); bar_wait(&s.bar_act[3], par); SYN_TICK(4)
        forward_async<4>(s, rank); bar_wait(&s.bar_act[4], par); SYN_TICK(5)
        { // 16 lanes per batch row, lane k < 12 owns logit k: head maxima and sums by xor-shuffles inside the half-warp
            const int b = t >> 4, k = t & 15;
            const bool isp = k < 9, isv = k >= 9 && k < 12;
            const float ninf = __uint_as_float(0xff800000u);
            const float z = k < 12 ? s.act[act_off(5) + b * 16 + k] : ninf;
            const float tgt = k < 12 ? s.target[b][k] : 0.0f;
            float mp = isp ? z : ninf, mv = isv ? z : ninf;
#pragma unroll
            for (int dlt = 8; dlt; dlt >>= 1) {
                mp = fmaxf(mp, __shfl_xor_sync(0xffffffffu, mp, dlt));
                mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, dlt));
            }
            const float mx = isp ? mp : mv;
            const float ex = k < 12 ? expf(z - mx) : 0.0f;
            float sep = isp ? ex : 0.f, sev = isv ? ex : 0.f, stp = isp ? tgt : 0.f, stv = isv ? tgt : 0.f;
#pragma unroll
            for (int dlt = 8; dlt; dlt >>= 1) {
                sep += __shfl_xor_sync(0xffffffffu, sep, dlt); sev += __shfl_xor_sync(0xffffffffu, sev, dlt);
                stp += __shfl_xor_sync(0xffffffffu, stp, dlt); stv += __shfl_xor_sync(0xffffffffu, stv, dlt);
            }
            const float se = isp ? sep : sev, st = isp ? stp : stv;
            const float lp = z - (mx + logf(se));
            float lk = (k < 12 && tgt > 0.f) ? tgt * (logf(tgt) - lp) : 0.f; // kl_div: xlogy(t, t) - t * input
            const float scale = (isp ? p.pw : p.vw) * inv_batch;
            const int lo = (int)rank * Own<4>::sl;
            if (k >= lo && k < lo + Own<4>::sl && k < 12) s.d[b * DSL + (k - lo)] = scale * (__fdividef(ex, se) * st - tgt);
            float lpol = isp ? lk : 0.f, lval = isv ? lk : 0.f;
#pragma unroll
            for (int dlt = 8; dlt; dlt >>= 1) {
                lpol += __shfl_xor_sync(0xffffffffu, lpol, dlt);
                lval += __shfl_xor_sync(0xffffffffu, lval, dlt);
            }
            if (k == 0) { s.loss_row[b][0] = lpol; s.loss_row[b][1] = lval; }
        }
        __syncthreads();
        if (rank == 0 && t < 2 && p.losses) {
            float tot = 0.f;
            for (int r = 0; r < B; ++r) tot += s.loss_row[r][t];
            const float part = tot * inv_batch;
            p.losses[(size_t)step * 2 + t] = k_micro == 0u ? part : p.losses[(size_t)step * 2 + t] + part;
        }
        SYN_TICK(6)
        backward_layer_async<4>(p, s, sc, rank, par, accum); SYN_TICK(7)
        backward_layer_async<3>(p, s, sc, rank, par, accum); SYN_TICK(8)
        backward_layer_async<2>(p, s, sc, rank, par, accum); SYN_TICK(9)
        backward_layer_async<1>(p, s, sc, rank, par, accum); SYN_TICK(10)
        backward_layer_async<0>(p, s, sc, rank, par, accum); SYN_TICK(11)
        cluster_barrier_relaxed(); // the next step's stores must not overtake this step's readers (execution order only)
        SYN_TICK(12)
        if (++k_micro == p.micro) { k_micro = 0u; ++step; }
    }
#undef SYN_TICK
    if (prof)
        for (int k = 0; k < 14; ++k) p.prof[k] = (unsigned long long)pc[k];
    sync_rows<0, false>(p, s, rank); sync_rows<1, false>(p, s, rank); sync_rows<2, false>(p, s, rank);
    sync_rows<3, false>(p, s, rank); sync_rows<4, false>(p, s, rank);
    cluster.sync(); // no CTA leaves while a peer could still address its shared memory
}

} // namespace trc
