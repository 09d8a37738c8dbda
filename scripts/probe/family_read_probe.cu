// family_read_probe.cu — how many "levels of selection" per second can HBM3e + L2 serve, per node layout?
//
// A level of select_best_child reads the records of one family (up to 9 children, contiguous) at an address that
// is effectively random in a footprint far larger than L2.  random_sector_probe.cu showed ~36.5 G random 32-byte
// sector reads/s; here the unit is the FAMILY and the question is what a layout costs:
//   mode 0  tpg2 records : 9 x 32 B contiguous, 32-byte aligned start (288 B span, 3-4 DRAM lines), LDG.256 each
//   mode 1  tpg3 records : 9 x 16 B at stride 32 B (same span, half of every sector used), LDG.128 each
//   mode 2  split layout : 9 x 16 B contiguous (144 B), 16-byte aligned start inside 256-byte family blocks
//   mode 3  one line     : 8 x 16 B contiguous, 128-byte aligned (a family that fits one DRAM line)
//   mode 4  one sector   : 1 x 32 B (random_sector_probe's unit, for reference)
// Every thread issues the loads of one family together, consumes them, and derives the next address from the data
// (a dependent chain like the tree walk).  Threads per SM = the kernels' occupancy.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o family_read_probe family_read_probe.cu && ./family_read_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned long long mix(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) probe(const uint4* __restrict__ buf, unsigned long long n16, int steps, unsigned long long* sink) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = tid * 0x9e3779b97f4a7c15ull;
    for (int s = 0; s < steps; ++s) {
        const unsigned long long r = mix(acc + (unsigned long long)s);
        if (MODE == 0) {
            const uint4* p = buf + (__umul64hi(r, n16 / 2 - 16) << 1); // 32-byte aligned
            unsigned long long q[9][4];
#pragma unroll
            for (int k = 0; k < 9; ++k)
                asm volatile("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(q[k][0]), "=l"(q[k][1]), "=l"(q[k][2]), "=l"(q[k][3]) : "l"(p + 2 * k));
#pragma unroll
            for (int k = 0; k < 9; ++k) acc += q[k][0] ^ q[k][1] ^ q[k][2] ^ q[k][3];
        } else {
            const uint4* p;
            int n = 9, stride = 1;
            if (MODE == 1) { p = buf + (__umul64hi(r, n16 / 2 - 16) << 1); stride = 2; }
            else if (MODE == 2) p = buf + (__umul64hi(r, n16 / 16 - 2) << 4) + (r & 7);       // 256-byte blocks, 16-byte aligned start, 144 B inside the block
            else if (MODE == 3) { p = buf + (__umul64hi(r, n16 / 8 - 2) << 3); n = 8; }        // 128-byte aligned
            else { p = buf + (__umul64hi(r, n16 / 2 - 2) << 1); n = 2; }                       // one 32-byte sector as two 16-byte loads
            uint4 q[9];
#pragma unroll
            for (int k = 0; k < 9; ++k)
                if (k < n) q[k] = p[k * stride];
#pragma unroll
            for (int k = 0; k < 9; ++k)
                if (k < n) acc += (unsigned long long)(q[k].x ^ q[k].y ^ q[k].z ^ q[k].w);
        }
    }
    if (acc == 0x1234567ull) *sink = acc;
}

template <int MODE>
static void run(const uint4* buf, double gib, int threads, int steps, unsigned long long* sink, int sms) {
    const unsigned long long n16 = (unsigned long long)(gib * (1ull << 30)) / 16;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<MODE><<<sms, threads>>>(buf, n16, steps / 4, sink); // warm-up
    cudaEventRecord(a);
    probe<MODE><<<sms, threads>>>(buf, n16, steps, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double fams = (double)sms * threads * steps;
    static const char* names[] = {"9x32B records (tpg2)", "9x16B stride 32 (tpg3)", "9x16B contiguous", "8x16B one line", "one 32B sector"};
    printf("footprint %5.1f GiB threads/SM %4d  %-24s: %6.2f G families/s  latency-per-step %6.0f ns\n", gib, threads, names[MODE], fams / ms / 1e6,
           (double)ms * 1e6 / steps);
}

int main(int argc, char** argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 32.0;
    const size_t bytes = (size_t)(gib * (1ull << 30));
    uint4* buf;
    unsigned long long* sink;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, bytes);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int threads : {128, 256, 640, 1024}) {
        const int steps = threads <= 256 ? 4096 : 2048;
        run<0>(buf, gib, threads, steps, sink, sms);
        run<1>(buf, gib, threads, steps, sink, sms);
        run<2>(buf, gib, threads, steps, sink, sms);
        run<3>(buf, gib, threads, steps, sink, sms);
        run<4>(buf, gib, threads, steps, sink, sms);
    }
    return 0;
}
