// sector_reach_probe.cu — WHY does random 32-byte sector traffic stop at ~36 G sectors/s (random_sector_probe.cu)?
//
// Three candidate limiters are separated here:
//   (1) DRAM: every 32-byte miss drags a whole 128-byte line out of HBM (ncu: 128 B of dram__bytes_read per
//       requested sector) — then the sector rate must not depend on the footprint once it exceeds L2;
//   (2) address translation: a warp-wide LDG.E.256 with 32 unrelated addresses touches 32 different 2 MB pages
//       — then the rate must rise when the footprint shrinks towards the TLB reach, and when the 32 lanes of
//       a warp stay inside ONE page (mode "warp-page");
//   (3) the request path SM -> L2 (per-sector packets) — then neither changes anything.
// Modes: 0 = every lane an independent random sector; 1 = the 32 lanes of a warp pick random sectors inside the
// same random 2 MB page; 2 = the 32 lanes pick random sectors inside the same random 8 MB window (what a warp of
// the thread-per-game kernel sees: 32 neighbouring 231 KB arenas); 3 = mode 0 but 128-byte aligned groups of 4
// sectors are read by 4 neighbouring lanes (full-line use).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sector_reach_probe sector_reach_probe.cu
//   ./sector_reach_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned long long mix(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <int K, int MODE>
__global__ void __launch_bounds__(1024, 1) probe(const uint4* __restrict__ buf, unsigned long long n_sectors, int steps, unsigned long long* sink) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long wid = tid >> 5;
    unsigned long long acc = 0;
    for (int s = 0; s < steps; ++s) {
        unsigned long long q[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned long long r = mix(tid * 0x9e3779b97f4a7c15ull + (unsigned long long)s * K + k);
            unsigned long long sector;
            if (MODE == 0) {
                sector = __umul64hi(r, n_sectors);
            } else if (MODE == 1 || MODE == 2) {
                const unsigned long long win = MODE == 1 ? (1ull << 16) : (1ull << 18); // sectors per window: 2 MB / 8 MB
                const unsigned long long rw = mix(wid * 0x9e3779b97f4a7c15ull + (unsigned long long)s * K + k + 77);
                sector = __umul64hi(rw, n_sectors / win) * win + (r & (win - 1));
            } else {
                const unsigned long long rq = mix((tid >> 2) * 0x9e3779b97f4a7c15ull + (unsigned long long)s * K + k + 99);
                sector = (__umul64hi(rq, n_sectors / 4) << 2) + (tid & 3);
            }
            const uint4* p = buf + 2 * sector;
            asm volatile("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(q[k][0]), "=l"(q[k][1]), "=l"(q[k][2]), "=l"(q[k][3]) : "l"(p));
        }
#pragma unroll
        for (int k = 0; k < K; ++k) acc += q[k][0] ^ q[k][1] ^ q[k][2] ^ q[k][3];
    }
    if (acc == 0x1234567ull) *sink = acc;
}

template <int K, int MODE>
static void run(const uint4* buf, double gib, int threads, int steps, unsigned long long* sink, int sms) {
    const unsigned long long n_sectors = (unsigned long long)(gib * (1ull << 30)) / 32;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<K, MODE><<<sms, threads>>>(buf, n_sectors, steps / 4, sink); // warm-up
    cudaEventRecord(a);
    probe<K, MODE><<<sms, threads>>>(buf, n_sectors, steps, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double sectors = (double)sms * threads * steps * K;
    static const char* names[] = {"independent", "warp-page 2MB", "warp-window 8MB", "line-quads"};
    printf("footprint %6.2f GiB  %-16s K=%d threads/SM %4d : %6.1f G sectors/s = %7.1f GB/s (%.1f ms)\n", gib, names[MODE], K, threads, sectors / ms / 1e6,
           sectors * 32 / ms / 1e6, ms);
}

int main(int argc, char** argv) {
    const double max_gib = argc > 1 ? atof(argv[1]) : 32.0;
    const size_t bytes = (size_t)(max_gib * (1ull << 30));
    uint4* buf;
    unsigned long long* sink;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, bytes);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (double gib : {0.25, 0.5, 1.0, 2.0, 4.0, 8.0, 16.0, 32.0}) {
        if (gib > max_gib) break;
        run<3, 0>(buf, gib, 512, 2048, sink, sms);
        run<3, 1>(buf, gib, 512, 2048, sink, sms);
        run<3, 2>(buf, gib, 512, 2048, sink, sms);
        run<3, 3>(buf, gib, 512, 2048, sink, sms);
    }
    return 0;
}
