/* syn_sampling.h — Normal / Gamma / Dirichlet samplers shared by the sm_100a kernels and the
 * CPU oracle.
 *
 * The reference draws root noise with rand_distr 0.4 `Dirichlet::new_with_size(alpha, k)`
 * from an UNSEEDED thread_rng (synthesis/src/mcts.rs:236,248-249) and its shipped FPU with
 * `Normal::new(1.0, 0.1)` from thread_rng as well (study-connect4/src/main.rs:43-47), so no
 * stream-level parity with the reference exists even CPU-vs-CPU ("parity unpinned").  What
 * is kept is the distribution: Dirichlet(alpha,...,alpha) as k independent Gamma(alpha, 1)
 * variates normalised by their f32 sum (the construction rand_distr uses), Gamma by
 * Marsaglia-Tsang (with the U^(1/alpha) boost for alpha < 1), Normal by Marsaglia's polar
 * method (rand_distr uses a ziggurat; same distribution, different stream use).
 *
 * Every operation is IEEE binary32 (+, *, /, sqrt) or syn_expf/syn_logf from syn_detmath.h, so
 * the device and the oracle produce identical bits from identical u32 streams; that is what
 * lets the parity tests stay bit-exact with noise switched on.
 *
 * R must provide `uint32_t next_u32()`.
 */
#ifndef SYN_SAMPLING_H
#define SYN_SAMPLING_H

#include "syn_detmath.h"

#if defined(__CUDACC__)
#define SYN_HDT __host__ __device__
#else
#define SYN_HDT
#endif

/* uniform in [0,1) with 24 bits */
SYN_HD float syn_u01(uint32_t v) { return (float)(v >> 8) * 5.9604644775390625e-8f; }
/* uniform in (0,1] with 24 bits */
SYN_HD float syn_u01_open0(uint32_t v) { return ((float)(v >> 8) + 1.0f) * 5.9604644775390625e-8f; }

template <class R>
SYN_HDT inline float syn_std_normal(R& r) {
    for (;;) {
        float u = 2.0f * syn_u01(r.next_u32()) - 1.0f;
        float v = 2.0f * syn_u01(r.next_u32()) - 1.0f;
        float s = u * u + v * v;
        if (s >= 1.0f || s == 0.0f) continue;
        float f = sqrtf((-2.0f * syn_logf(s)) / s);
        return u * f;
    }
}

template <class R>
SYN_HDT inline float syn_normal(R& r, float mean, float std) {
    return mean + std * syn_std_normal(r);
}

template <class R>
SYN_HDT inline float syn_gamma(R& r, float alpha) {
    float boost = 1.0f;
    float a = alpha;
    if (alpha < 1.0f) {
        float u = syn_u01_open0(r.next_u32());
        boost = syn_expf(syn_logf(u) / alpha);
        a = alpha + 1.0f;
    }
    float d = a - (1.0f / 3.0f);
    float c = 1.0f / sqrtf(9.0f * d);
    for (;;) {
        float x = syn_std_normal(r);
        float v = 1.0f + c * x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        float u = syn_u01_open0(r.next_u32());
        float x2 = x * x;
        float rhs = (0.5f * x2 + d) - d * v + d * syn_logf(v);
        if (syn_logf(u) < rhs) return (d * v) * boost;
    }
}

/* out[0..k) ~ Dirichlet(alpha, ..., alpha) */
template <class R>
SYN_HDT inline void syn_dirichlet(R& r, float alpha, int k, float* out) {
    float total = 0.0f;
    for (int i = 0; i < k; ++i) {
        out[i] = syn_gamma(r, alpha);
        total += out[i];
    }
    for (int i = 0; i < k; ++i) out[i] = out[i] / total;
}

#endif /* SYN_SAMPLING_H */
