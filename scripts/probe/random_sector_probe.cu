// random_sector_probe.cu — what can HBM3e deliver for the tree kernels' access pattern?
//
// The self-play kernels read 32-byte node records at data-dependent addresses spread over tens of
// GB (one arena per game).  This probe measures the DRAM bandwidth a B200 sustains for exactly that:
// every thread issues K independent 32-byte loads (LDG.E.256) per step at pseudo-random sector
// addresses inside a buffer far larger than L2, over a grid of one persistent CTA per SM.  The
// result is the ceiling the kernels' measured DRAM traffic should be compared with.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o random_sector_probe random_sector_probe.cu
//   ./random_sector_probe [GiB=32] [CTAs=#SMs] [quick=0]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned long long mix(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <int K, bool CONTIG> // K independent loads in flight per thread; CONTIG: the K sectors are consecutive (one child block)
__global__ void __launch_bounds__(1024, 1) probe(const uint4* __restrict__ buf, unsigned long long n_sectors, int steps, unsigned long long* sink) {
    unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    for (int s = 0; s < steps; ++s) {
        unsigned long long q[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            unsigned long long base = CONTIG ? mix(tid * 0x9e3779b97f4a7c15ull + (unsigned long long)s) % (n_sectors - K) + k
                                             : mix(tid * 0x9e3779b97f4a7c15ull + (unsigned long long)s * K + k) % n_sectors;
            const uint4* p = buf + 2 * base;
            asm volatile("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(q[k][0]), "=l"(q[k][1]), "=l"(q[k][2]), "=l"(q[k][3]) : "l"(p));
        }
#pragma unroll
        for (int k = 0; k < K; ++k) acc += q[k][0] ^ q[k][1] ^ q[k][2] ^ q[k][3];
    }
    if (acc == 0x1234567ull) *sink = acc;
}

template <int K, bool CONTIG>
static void run(const uint4* buf, unsigned long long n_sectors, int threads, int steps, unsigned long long* sink, int sms) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<K, CONTIG><<<sms, threads>>>(buf, n_sectors, steps / 4, sink); // warm-up
    cudaEventRecord(a);
    probe<K, CONTIG><<<sms, threads>>>(buf, n_sectors, steps, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    double sectors = (double)sms * threads * steps * K;
    printf("threads/SM %4d  %s loads in flight/thread %d : %7.1f G sectors/s = %7.1f GB/s of 32-byte records (%.1f ms)\n", threads, CONTIG ? "consecutive" : "independent", K,
           sectors / ms / 1e6, sectors * 32 / ms / 1e6, ms);
}

int main(int argc, char** argv) {
    double gib = argc > 1 ? atof(argv[1]) : 32.0;
    size_t bytes = (size_t)(gib * (1ull << 30));
    uint4* buf;
    unsigned long long* sink;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, bytes);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (argc > 2 && atoi(argv[2]) > 0) sms = atoi(argv[2]);
    const bool quick = argc > 3 && atoi(argv[3]) != 0;
    unsigned long long n_sectors = bytes / 32;
    printf("random 32-byte sector reads over %.0f GiB, %d CTAs\n", gib, sms);
    if (quick) { // one launch pair per pattern: what ncu is pointed at
        run<3, false>(buf, n_sectors, 512, 2048, sink, sms);
        run<9, true>(buf, n_sectors, 512, 1024, sink, sms);
        return 0;
    }
    for (int threads : {256, 512, 1024}) {
        run<1, false>(buf, n_sectors, threads, 4096, sink, sms);
        run<3, false>(buf, n_sectors, threads, 2048, sink, sms);
        run<9, false>(buf, n_sectors, threads, 1024, sink, sms);
        run<3, true>(buf, n_sectors, threads, 2048, sink, sms);
        run<9, true>(buf, n_sectors, threads, 1024, sink, sms);
    }
    return 0;
}
