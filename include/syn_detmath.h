/* syn_detmath.h — deterministic single-precision expf / logf shared by the sm_100a kernels
 * and by the CPU oracle.
 *
 * Why this exists: the reference computes child priors with `f32::exp`
 * (synthesis/src/mcts.rs:418) and the UCT bonus with `f32::ln` (mcts.rs:364,
 * evaluator.rs:420).  Rust lowers both to the platform libm, CUDA has its own
 * implementation, and the two differ in the last ulp on a small fraction of inputs.
 * MCTS selection compares `q + u` with a strict `>` (mcts.rs:335), so a one-ulp
 * difference can flip an argmax and the trees diverge.  Bit-exact visit counts on two
 * machines therefore need ONE definition of exp and ln.  These two functions are that
 * definition: every operation is an IEEE-754 binary32 add, multiply, divide, or a
 * correctly-rounded fused multiply-add spelled explicitly as fmaf(), so the result is a
 * pure function of the input bits on any conforming host or device — provided the
 * translation unit is compiled without FP contraction and without fast-math
 * (nvcc: -fmad=false, default -prec-div/-prec-sqrt; gcc: -ffp-contract=off).
 *
 * Accuracy (checked in tests/test_shared_math_gap.py against glibc in double): <= 1 ulp for
 * expf on [-104, 88.7], <= 1 ulp for logf on positive normals.
 *
 * The polynomial coefficients are the classic Cephes single-precision ones.
 */
#ifndef SYN_DETMATH_H
#define SYN_DETMATH_H

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define SYN_HD __host__ __device__ __forceinline__
#else
#define SYN_HD static inline
#endif

SYN_HD uint32_t syn_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
SYN_HD float syn_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

/* 2^k for k in [-126, 127], exact. */
SYN_HD float syn_pow2i(int k) { return syn_u2f((uint32_t)(k + 127) << 23); }

SYN_HD float syn_expf(float x) {
    if (x != x) return x;
    if (x > 88.72283f) return syn_u2f(0x7f800000u);
    if (x < -103.97208f) return 0.0f;
    float t = x * 1.44269504088896341f;
    float n = floorf(t + 0.5f);
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float y = fmaf(p, r2, r);
    y = y + 1.0f;
    int k = (int)n;
    int k1 = k / 2;
    int k2 = k - k1;
    return (y * syn_pow2i(k1)) * syn_pow2i(k2);
}

SYN_HD float syn_logf(float x) {
    if (x != x) return x;
    if (x < 0.0f) return syn_u2f(0x7fc00000u);
    if (x == 0.0f) return syn_u2f(0xff800000u);
    uint32_t u = syn_f2u(x);
    if (u == 0x7f800000u) return x;
    int e = 0;
    if (u < 0x00800000u) { /* subnormal: scale by 2^23 (exact) */
        x = x * 8388608.0f;
        u = syn_f2u(x);
        e = -23;
    }
    /* x = m * 2^e with m in [0.5, 1) */
    e += (int)(u >> 23) - 126;
    float m = syn_u2f((u & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) {
        e -= 1;
        m = (m + m) - 1.0f;
    } else {
        m = m - 1.0f;
    }
    float z = m * m;
    float p = 7.0376836292e-2f;
    p = fmaf(p, m, -1.1514610310e-1f);
    p = fmaf(p, m, 1.1676998740e-1f);
    p = fmaf(p, m, -1.2420140846e-1f);
    p = fmaf(p, m, 1.4249322787e-1f);
    p = fmaf(p, m, -1.6668057665e-1f);
    p = fmaf(p, m, 2.0000714765e-1f);
    p = fmaf(p, m, -2.4999993993e-1f);
    p = fmaf(p, m, 3.3333331174e-1f);
    float fe = (float)e;
    float y = (m * z) * p;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(z, -0.5f, y);
    float r = m + y;
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

#endif /* SYN_DETMATH_H */
