// devport.cuh — lets the device-side TREE code (c4.cuh, treedefs.cuh, tpg4_tree.cuh) compile as plain C++.
//
// With SYN_HOST_EMU defined (tests/emu/: g++ -ffp-contract=off, never nvcc) the CUDA qualifiers vanish and the
// handful of intrinsics the tree code uses become their IEEE / integer definitions, so that the SAME source that
// runs in the kernels can be stepped on a CPU and compared with the oracle node for node.  This is test
// plumbing for the product's own source; nothing in the library is built with SYN_HOST_EMU and the product has
// no CPU path.
#pragma once
#ifdef SYN_HOST_EMU
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
struct uint4 { uint32_t x, y, z, w; };
struct float2 { float x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(uint64_t x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif
