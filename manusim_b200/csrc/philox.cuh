// Philox4x32-10 counter-based generator + the six manusim distributions as device functions.
//
// Variates are keyed (replication key, stream id, draw index[, rejection iteration]) so that a draw never
// depends on how many uniforms earlier draws consumed (BASELINE.json north_star "Random variates").
// Distributions follow DistributionGenerator (/root/reference/src/manusim/engine/utils.py:42-102):
// scalar draws are clamped with max(0, float32(v)); batch sums of `count` iid samples use the closed
// forms Gamma(count*k, theta) / Gamma(count, mean) for gamma/erlang/expo and a loop otherwise.
#pragma once
#include <stdint.h>

namespace msim {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    // the ten rounds stay a loop on the device: the simulation kernel is instruction-fetch bound and every inlined copy of
    // the unrolled rounds is ~1.3 KB of rarely executed code (profiles/r02_ab_variants_k_*.jsonl)
#ifdef __CUDA_ARCH__
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 24-bit uniform strictly inside (0,1)
__host__ __device__ __forceinline__ float u01(uint32_t x) {
    return (float)(x >> 8) * 5.9604644775390625e-08f + 2.98023223876953125e-08f;
}

struct PhiloxStream {
    uint32_t k0, k1;     // replication key
    uint32_t stream;     // 0=A rnd_product, 1=B rnd_process scalar, 2=C rnd_process batch, 3=D rnd_breakdown
    uint32_t draw;       // draw index on that stream
};

#ifdef __CUDACC__
// Every draw index of every stream owns ONE Philox block, turned into a standard normal x (Box-Muller on words
// 0,1) and an independent uniform u (word 2).  All six distributions are functions of that (x,u) pair, so a draw
// has the same value whether its pair was precomputed by the whole warp (bulk refill in the kernel's cooperative
// phase) or generated on the spot by lane 0 (slow path: boot, breakdown stream, buffer miss).
struct XU { float x, u; };
constexpr int kXuBuf = 16;          // precomputed pairs per stream and replication
constexpr uint32_t kXuRefillAt = 4; // refill when this many (or fewer) remain

__device__ __forceinline__ XU xu_from_words(const uint32_t w[4]) {
    XU r;
    float m = sqrtf(-2.0f * __logf(u01(w[0])));
    r.x = m * cospif(2.0f * u01(w[1]));
    r.u = u01(w[2]);
    return r;
}

// one out-of-line copy per kernel: the simulation kernel is instruction-cache bound, not ALU bound
// (static: this header is included by two translation units of libmsim.so)
static __device__ __noinline__ XU raw_xu(uint32_t k0, uint32_t k1, uint32_t stream, uint32_t draw, uint32_t it) {
    uint32_t w[4];
    philox4x32_10(draw, stream, it, 0u, k0, k1, w);
    return xu_from_words(w);
}

// gamma shape < 1: Gamma(k) = Gamma(k+1) * U^(1/k)
static __device__ __noinline__ float gamma_boost(const PhiloxStream s, float k) {
    uint32_t w[4];
    philox4x32_10(s.draw, s.stream, 0xFFFFu, 1u, s.k0, s.k1, w);
    return __powf(u01(w[0]), 1.0f / k);
}

// Marsaglia-Tsang rejection iterations 1.. (iteration 0 is the precomputed pair); d = kk-1/3, c = 1/sqrt(9d)
static __device__ __noinline__ float gamma_retry(const PhiloxStream s, float d, float c) {
    float out = d;
    for (uint32_t it = 1; it < 64u; ++it) {
        XU r = raw_xu(s.k0, s.k1, s.stream, s.draw, it);
        float v = 1.0f + c * r.x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        float x2 = r.x * r.x;
        out = d * v;
        if (r.u < 1.0f - 0.0331f * x2 * x2) break;
        if (__logf(r.u) < 0.5f * x2 + d * (1.0f - v + __logf(v))) break;
    }
    return out;
}

__device__ __forceinline__ float gamma_from_xu(const PhiloxStream& s, float k, float theta, float d, float c, XU r) {
    float g;
    float v = 1.0f + c * r.x;
    float v3 = v * v * v;
    float x2 = r.x * r.x;
    bool ok = v > 0.0f;
    if (ok) ok = (r.u < 1.0f - 0.0331f * x2 * x2) || (__logf(r.u) < 0.5f * x2 + d * (1.0f - v3 + __logf(v3)));
    g = ok ? d * v3 : gamma_retry(s, d, c);
    if (k < 1.0f) g *= gamma_boost(s, k);
    return g * theta;
}

// value of one sample BEFORE the max(0, .) clamp; (d0,d1) derived parameters, (md,mc) Marsaglia-Tsang constants
__device__ __forceinline__ float sample_from_xu(const PhiloxStream& s, int kind, float d0, float d1, float md, float mc,
                                                XU r) {
    switch (kind) {
    case 1: return d0 + (d1 - d0) * r.u;                     // uniform(a,b)
    case 2: case 3: return gamma_from_xu(s, d0, d1, md, mc, r);
    case 4: return -d0 * __logf(r.u);                        // expo(mean)
    case 5: return d0 + d1 * r.x;                            // normal(mu, sigma)
    default: return d0;
    }
}

// sums of count > 1 samples (utils.py:83-98): closed forms for gamma/erlang/expo, loops otherwise
static __device__ __noinline__ float batch_many(const PhiloxStream s, int kind, float d0, float d1, int count, XU r) {
    if (kind == 2 || kind == 3 || kind == 4) {
        float k = kind == 4 ? (float)count : d0 * (float)count;
        float th = kind == 4 ? d0 : d1;
        float kk = k < 1.0f ? k + 1.0f : k;
        float d = kk - (1.0f / 3.0f);
        return gamma_from_xu(s, k, th, d, rsqrtf(9.0f * d), r);
    }
    float t = 0.0f;
    for (int i = 0; i < count; ++i) {
        XU q = i == 0 ? r : raw_xu(s.k0, s.k1, s.stream, s.draw, 16u + (uint32_t)i);
        float v = sample_from_xu(s, kind, d0, d1, 0.0f, 0.0f, q);
        t += (kind == 5 && !(v > 0.0f)) ? 0.0f : v;          // normal: per-sample float32 clamp
    }
    return t;
}
#endif

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// replication key from the experiment key and the GLOBAL replication index (shard-invariant)
__host__ __device__ __forceinline__ uint64_t rep_key(uint64_t seed, uint64_t global_rep) {
    return splitmix64(seed ^ splitmix64(global_rep + 0x632BE59BD9B4E019ull));
}

} // namespace msim
