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
#pragma unroll
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
// one out-of-line copy per kernel: the simulation kernel is instruction-cache bound, not ALU bound
__device__ __noinline__ void philox_nl(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                       uint32_t out[4]) {
    philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}

__device__ __forceinline__ float normal01(const uint32_t w[4]) {
    // Box-Muller on two of the four words
    float r = sqrtf(-2.0f * __logf(u01(w[0])));
    return r * cospif(2.0f * u01(w[1]));
}

// Marsaglia-Tsang; iteration index is part of the counter so rejections stay reproducible
__device__ __noinline__ float gamma_mt(const PhiloxStream s, float k, float theta) {
    float boost = 1.0f;
    float kk = k;
    uint32_t w[4];
    if (k < 1.0f) {
        philox_nl(s.draw, s.stream, 0xFFFFu, 1u, s.k0, s.k1, w);
        boost = __powf(u01(w[0]), 1.0f / k);
        kk = k + 1.0f;
    }
    const float d = kk - (1.0f / 3.0f);
    const float c = rsqrtf(9.0f * d);
    float out = d;
    for (uint32_t it = 0; it < 64u; ++it) {
        philox_nl(s.draw, s.stream, it, 0u, s.k0, s.k1, w);
        float x = normal01(w);
        float v = 1.0f + c * x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        float u = u01(w[2]);
        float x2 = x * x;
        out = d * v;
        if (u < 1.0f - 0.0331f * x2 * x2) break;
        if (__logf(u) < 0.5f * x2 + d * (1.0f - v + __logf(v))) break;
    }
    return out * theta * boost;
}

// value of one sample BEFORE the max(0, .) clamp; kind/d0/d1 are the derived parameters
__device__ __forceinline__ float sample_raw(const PhiloxStream& s, int kind, float d0, float d1, uint32_t sub) {
    uint32_t w[4];
    switch (kind) {
    case 1: // uniform(a,b)
        philox_nl(s.draw, s.stream, sub, 2u, s.k0, s.k1, w);
        return d0 + (d1 - d0) * u01(w[0]);
    case 2: case 3:
        return gamma_mt(s, d0, d1);
    case 4: // expo(mean)
        philox_nl(s.draw, s.stream, sub, 2u, s.k0, s.k1, w);
        return -d0 * __logf(u01(w[0]));
    case 5: // normal(mu, sigma)
        philox_nl(s.draw, s.stream, sub, 2u, s.k0, s.k1, w);
        return d0 + d1 * normal01(w);
    default:
        return d0;
    }
}

// DistributionGenerator.random_number (utils.py:42-67): max(0, np.float32(v))
__device__ __noinline__ float draw_scalar_philox(const PhiloxStream s, int kind, float d0, float d1) {
    float v = sample_raw(s, kind, d0, d1, 0u);
    return v > 0.0f ? v : 0.0f;
}

// DistributionGenerator.sum_random_numbers_batch (utils.py:69-102)
__device__ __noinline__ double draw_batch_philox(const PhiloxStream s, int kind, float d0, float d1, double raw0,
                                                    int count) {
    if (count <= 0) return 0.0;
    switch (kind) {
    case 0: return (double)count * raw0;
    case 2: case 3: return (double)gamma_mt(s, d0 * (float)count, d1);
    case 4: return (count == 1) ? (double)sample_raw(s, 4, d0, d1, 0u) : (double)gamma_mt(s, (float)count, d0);
    case 1: {
        float t = 0.0f;
        for (int i = 0; i < count; ++i) t += sample_raw(s, 1, d0, d1, 16u + (uint32_t)i);
        return (double)t;
    }
    default: { // normal: per-sample float32 clamp (utils.py:97-98)
        float t = 0.0f;
        for (int i = 0; i < count; ++i) {
            float v = sample_raw(s, 5, d0, d1, 16u + (uint32_t)i);
            t += v > 0.0f ? v : 0.0f;
        }
        return (double)t;
    }
    }
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
