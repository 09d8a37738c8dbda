// dedup.cuh — ReplayBuffer::deduplicate on the device.
//
// Replaces synthesis/src/data.rs:196-235 of the reference: rows of the experience buffer that hold
// the same game position (HashMap<G, StateStatistics>, key = the Connect4 value, i.e. its two
// bitboards — height and player are functions of them, study-connect4/src/connect4.rs:108-121) are
// merged into ONE training row whose policy and value targets are the arithmetic mean over the
// group: `sum_pi[j] += pis[i][j]` for i in BUFFER ORDER (data.rs:199-215), then `sum / num as f32`
// (data.rs:220-229).  f32 addition does not associate, so the sums here run in ascending row index
// too, one component per lane, and the results are bit-identical to the reference's.
//
// The reference emits the groups in HashMap iteration order (random per process); here the order
// is that of each position's FIRST occurrence in the buffer — deterministic, and equal as a set.
//
// Pipeline (all HBM-streaming except the hash probes):
//   insert   : open-addressing table of row indices, linear probing, one 32-bit CAS per new key;
//              the table entry names a row that holds the key, slot_first[] = min row of the group
//   rep      : every row learns its group's first row (`rep`) — a 32-bit group key that is already
//              in first-occurrence order
//   sort     : stable LSD radix sort of (rep, row) pairs, 8 bits per pass, ceil(log2 n / 8) passes;
//              per-warp match_any ranking keeps equal keys in row order
//   heads    : group boundaries -> exclusive scan -> group start table
//   gather   : policy/value targets permuted into sorted order (the only pass of random reads)
//   reduce   : 16 lanes per group (9 policy + 3 value components, sequential f32 sums in row
//              order); groups above BIG rows are streamed by one CTA each through shared memory
//   features : Game::features() of the group's position, synthesised from the bitboards
#pragma once
#include <stdint.h>

#include "c4.cuh"

namespace dd {

constexpr uint32_t EMPTY = 0xffffffffu;
constexpr int T = 256;            // threads per CTA everywhere in this file
constexpr int RS_K = 8;           // 32-key chunks per warp and tile
constexpr int RS_TILE = T * RS_K; // keys per CTA and pass
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = T * SCAN_IPT;
constexpr uint32_t BIG = 2048;    // groups with more rows than this get a CTA of their own

__device__ __forceinline__ uint64_t hash_key(uint64_t my, uint64_t op) {
    uint64_t x = (my * 0x9E3779B97F4A7C15ull) ^ ((op + 0xD1B54A32D192ED03ull) * 0xBF58476D1CE4E5B9ull);
    x ^= x >> 32;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 29;
    return x;
}

__global__ void fill_kernel(uint32_t* a, size_t n, uint32_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) a[i] = v;
}

// HashMap::entry(game) (data.rs:200-208): find or claim the slot of row i's position.
__global__ void insert_kernel(const uint64_t* __restrict__ my, const uint64_t* __restrict__ op, uint32_t n, uint32_t* slot_row,
                              uint32_t* slot_first, uint32_t mask, uint32_t* __restrict__ slot_of) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t m = my[i], o = op[i];
    uint32_t s = (uint32_t)hash_key(m, o) & mask;
    for (;;) {
        uint32_t cur = *(volatile uint32_t*)(slot_row + s);
        if (cur == EMPTY) {
            cur = atomicCAS(slot_row + s, EMPTY, i);
            if (cur == EMPTY) cur = i;
        }
        if (cur == i || (my[cur] == m && op[cur] == o)) {
            if (*(volatile uint32_t*)(slot_first + s) > i) atomicMin(slot_first + s, i);
            slot_of[i] = s;
            return;
        }
        s = (s + 1u) & mask;
    }
}

// rep[i] = first row of i's group; written over slot_of.
__global__ void rep_kernel(uint32_t* __restrict__ slot_of_then_rep, const uint32_t* __restrict__ slot_first, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_of_then_rep[i] = slot_first[slot_of_then_rep[i]];
}

// ---------------------------------------------------------------- exclusive scan (u32), reduce-then-scan
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* wsum /*[8] shared*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) {
        uint32_t s = wsum[w];
        if (w < warp) woff += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return woff + inc - v;
}

__global__ void __launch_bounds__(T) scan_reduce_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t wsum[T / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k)
        if (base + k < n) s += in[base + k];
    uint32_t tot;
    block_exclusive_scan(s, &tot, wsum);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// out[i] = offs[block] + exclusive prefix inside the tile; in == out is allowed.  `total` (optional) gets the grand total.
__global__ void __launch_bounds__(T) scan_down_kernel(const uint32_t* in, uint32_t* out, uint32_t n, const uint32_t* __restrict__ offs,
                                                      uint32_t* __restrict__ total) {
    __shared__ uint32_t wsum[T / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    uint32_t x[SCAN_IPT], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        x[k] = base + k < n ? in[base + k] : 0u;
        s += x[k];
    }
    uint32_t tot;
    uint32_t run = block_exclusive_scan(s, &tot, wsum) + (offs ? offs[blockIdx.x] : 0u);
    if (total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total = (offs ? offs[blockIdx.x] : 0u) + tot;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        if (base + k < n) out[base + k] = run;
        run += x[k];
    }
}

// ---------------------------------------------------------------- stable LSD radix sort, 8 bits per pass
// Tile layout: warp w of the CTA owns keys [tile + w*RS_K*32, tile + (w+1)*RS_K*32), chunk k = 32 consecutive keys.
__global__ void __launch_bounds__(T) rs_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ hist,
                                                    uint32_t nblk) {
    __shared__ uint32_t cnt[256];
    cnt[threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * RS_TILE + warp * (RS_K * 32) + lane;
#pragma unroll
    for (int k = 0; k < RS_K; ++k) {
        const uint32_t idx = base + k * 32;
        const bool valid = idx < n;
        const uint32_t d = valid ? ((keys[idx] >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == __ffs(peers) - 1) atomicAdd(&cnt[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    hist[threadIdx.x * nblk + blockIdx.x] = cnt[threadIdx.x]; // digit-major: the flat exclusive scan is the scatter base
}

// vals_in == nullptr: the values are the row indices themselves (first pass).
__global__ void __launch_bounds__(T) rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                       uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                       const uint32_t* __restrict__ hist_scanned, uint32_t nblk) {
    __shared__ uint32_t wcnt[T / 32][256];
    for (int w = 0; w < T / 32; ++w) wcnt[w][threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * RS_TILE + warp * (RS_K * 32) + lane;
    uint32_t key[RS_K], lr[RS_K];
#pragma unroll
    for (int k = 0; k < RS_K; ++k) {
        const uint32_t idx = base + k * 32;
        const bool valid = idx < n;
        key[k] = valid ? keys_in[idx] : 0u;
        const uint32_t d = valid ? ((key[k] >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0u;
        if (valid && lane == leader) {
            old = wcnt[warp][d];
            wcnt[warp][d] = old + (uint32_t)__popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        lr[k] = old + (uint32_t)__popc(peers & ((1u << lane) - 1u)); // rank among equal digits, in key order
        __syncwarp();
    }
    __syncthreads();
    { // digit = threadIdx.x: bases of the CTA's warps, in warp order
        uint32_t run = hist_scanned[threadIdx.x * nblk + blockIdx.x];
#pragma unroll
        for (int w = 0; w < T / 32; ++w) {
            uint32_t c = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_K; ++k) {
        const uint32_t idx = base + k * 32;
        if (idx < n) {
            const uint32_t dst = wcnt[warp][(key[k] >> shift) & 255u] + lr[k];
            keys_out[dst] = key[k];
            vals_out[dst] = vals_in ? vals_in[idx] : idx;
        }
    }
}

// ---------------------------------------------------------------- groups
__global__ void heads_kernel(const uint32_t* __restrict__ rep_sorted, uint32_t n, uint32_t* __restrict__ flags) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) flags[j] = (j == 0u || rep_sorted[j] != rep_sorted[j - 1u]) ? 1u : 0u;
}

__global__ void starts_kernel(const uint32_t* __restrict__ rep_sorted, const uint32_t* __restrict__ gid, uint32_t n,
                              const uint32_t* __restrict__ n_groups, uint32_t* __restrict__ gstart) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (j == 0u || rep_sorted[j] != rep_sorted[j - 1u]) gstart[gid[j]] = j;
    if (j == n - 1u) gstart[*n_groups] = n;
}

// Targets in sorted order: pv[j][0..9) = pis[row_j], pv[j][9..12) = vs[row_j].  The one pass of random reads; both
// reduce kernels then stream contiguous memory (a big group's rows are spread over the whole buffer, and chasing them
// from a single CTA is bound by TLB misses and DRAM latency, not bandwidth).
__global__ void __launch_bounds__(T) gather_kernel(const uint32_t* __restrict__ rows_sorted, uint32_t n, const float* __restrict__ pis,
                                                   const float* __restrict__ vs, float* __restrict__ pv) {
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, c = threadIdx.x & 15u;
    if (j >= n || c >= 12u) return;
    const uint32_t idx = rows_sorted[j];
    pv[(size_t)j * 12 + c] = c < 9u ? pis[(size_t)idx * 9 + c] : vs[(size_t)idx * 3 + (c - 9u)];
}

struct Out {
    uint64_t* my_bb; // [U]
    uint64_t* op_bb; // [U]
    uint32_t* num;   // [U]       StateStatistics::num
    float* states;   // [U][63]   StateStatistics::state (Game::features of the position)
    float* pis;      // [U][9]    sum_pi / num
    float* vs;       // [U][3]    sum_v / num
};

__device__ __forceinline__ void emit_header(const Out& o, uint32_t g, uint64_t m, uint64_t p, uint32_t cnt, int lane, int lanes) {
    if (lane == 0) {
        if (o.my_bb) o.my_bb[g] = m;
        if (o.op_bb) o.op_bb[g] = p;
        if (o.num) o.num[g] = cnt;
    }
    if (o.states)
        for (int i = lane; i < 63; i += lanes) o.states[(size_t)g * 63 + i] = c4::feature(m, p, i);
}

// 16 lanes per group; lane c < 9 owns sum_pi[c], lanes 9..11 own sum_v[c - 9].  pv = targets in sorted order.
__global__ void __launch_bounds__(T) reduce_kernel(const uint32_t* __restrict__ rows_sorted, const uint32_t* __restrict__ gstart,
                                                   const uint32_t* __restrict__ n_groups, const uint64_t* __restrict__ my,
                                                   const uint64_t* __restrict__ op, const float* __restrict__ pv, Out o,
                                                   uint32_t* __restrict__ big_list, uint32_t* __restrict__ big_count) {
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int c = threadIdx.x & 15;
    if (g >= *n_groups) return;
    const uint32_t s = gstart[g], e = gstart[g + 1u], cnt = e - s;
    if (cnt > BIG) {
        if (c == 0) big_list[atomicAdd(big_count, 1u)] = g;
        return;
    }
    float acc = 0.0f;
    if (c < 12) {
        const float* src = pv + (size_t)s * 12 + c;
        uint32_t r = 0;
        for (; r + 8u <= cnt; r += 8u) {
            float x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = src[(size_t)(r + u) * 12];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, x[u]);
        }
        for (; r < cnt; ++r) acc = __fadd_rn(acc, src[(size_t)r * 12]);
    }
    const float avg = __fdiv_rn(acc, (float)cnt);
    if (c < 9) { if (o.pis) o.pis[(size_t)g * 9 + c] = avg; }
    else if (c < 12) { if (o.vs) o.vs[(size_t)g * 3 + (c - 9)] = avg; }
    const uint32_t first_row = rows_sorted[s];
    emit_header(o, g, my[first_row], op[first_row], cnt, c, 16);
}

// One CTA of 512 threads per big group: warps 1..15 copy the next BIG_CHUNK rows of the sorted targets (contiguous: 128-bit
// loads, one round trip per chunk) into shared memory while lanes 0..11 of warp 0 add the previous chunk in row order.
// The chain of dependent f32 additions (4 cycles per row) is the critical path of the whole call.
constexpr int BIG_T = 512;
constexpr int BIG_STAGERS = BIG_T - 32;
constexpr int BIG_CHUNK = 2 * BIG_STAGERS; // 960 rows, 2 x 45 KB of shared memory
constexpr size_t BIG_SMEM = 2 * (size_t)BIG_CHUNK * 12 * sizeof(float);

__global__ void __launch_bounds__(BIG_T) reduce_big_kernel(const uint32_t* __restrict__ rows_sorted, const uint32_t* __restrict__ gstart,
                                                           const uint32_t* __restrict__ big_list, const uint64_t* __restrict__ my,
                                                           const uint64_t* __restrict__ op, const float* __restrict__ pv, Out o) {
    extern __shared__ __align__(16) float big_buf[]; // [2][BIG_CHUNK * 12]
    const uint32_t g = big_list[blockIdx.x];
    const uint32_t s = gstart[g], e = gstart[g + 1u], cnt = e - s;
    const uint32_t chunks = (cnt + BIG_CHUNK - 1) / BIG_CHUNK;
    float acc = 0.0f;
    auto stage = [&](uint32_t k) { // threads 32 .. BIG_T-1; a row is 48 bytes, so every chunk starts 16-byte aligned
        float4* b = reinterpret_cast<float4*>(big_buf + (k & 1u) * (BIG_CHUNK * 12));
        const uint32_t j0 = s + k * BIG_CHUNK, m = min((uint32_t)BIG_CHUNK, e - j0);
        const float4* src = reinterpret_cast<const float4*>(pv + (size_t)j0 * 12);
        const uint32_t n4 = m * 3u, t0 = threadIdx.x - 32u;
        float4 x[6];
#pragma unroll
        for (uint32_t u = 0; u < 6u; ++u) x[u] = t0 + u * BIG_STAGERS < n4 ? src[t0 + u * BIG_STAGERS] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (uint32_t u = 0; u < 6u; ++u)
            if (t0 + u * BIG_STAGERS < n4) b[t0 + u * BIG_STAGERS] = x[u];
    };
    if (threadIdx.x >= 32) stage(0);
    __syncthreads();
    for (uint32_t k = 0; k < chunks; ++k) {
        if (threadIdx.x >= 32) {
            if (k + 1u < chunks) stage(k + 1u);
        } else if (threadIdx.x < 12) {
            const float* b = big_buf + (k & 1u) * (BIG_CHUNK * 12) + threadIdx.x;
            const uint32_t m = min((uint32_t)BIG_CHUNK, e - (s + k * BIG_CHUNK));
            // the dependent additions (4 cycles each) are the critical path: the next sixteen operands are always in
            // registers before the current sixteen have been added
            uint32_t r = 0;
            float x[16], y[16];
            if (m >= 16u) {
#pragma unroll
                for (int u = 0; u < 16; ++u) x[u] = b[u * 12u];
                for (; r + 32u <= m; r += 16u) {
#pragma unroll
                    for (int u = 0; u < 16; ++u) y[u] = b[(r + 16u + u) * 12u];
#pragma unroll
                    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, x[u]);
#pragma unroll
                    for (int u = 0; u < 16; ++u) x[u] = y[u];
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, x[u]);
                r += 16u;
            }
            for (; r < m; ++r) acc = __fadd_rn(acc, b[r * 12u]);
        }
        __syncthreads();
    }
    const int c = threadIdx.x;
    if (c < 12) {
        const float avg = __fdiv_rn(acc, (float)cnt);
        if (c < 9) { if (o.pis) o.pis[(size_t)g * 9 + c] = avg; }
        else { if (o.vs) o.vs[(size_t)g * 3 + (c - 9)] = avg; }
    }
    const uint32_t first_row = rows_sorted[s];
    emit_header(o, g, my[first_row], op[first_row], cnt, threadIdx.x, BIG_T);
}

} // namespace dd
