// ORACLE — TEST INFRASTRUCTURE ONLY (see jl_math.hpp header).
//
// CPU restatement of the Julia stdlib `Random` pieces the reference's envs/policies
// consume (Julia 1.10: stdlib/Random/src/Xoshiro.jl, XoshiroSimd.jl, generation.jl).
// `Random` is NOT under /root/reference (Julia stdlib, version pinned only by
// .devcontainer/Dockerfile:1 `julia:1.10`) -> PARITY UNPINNED; restated from the
// published xoshiro256++ algorithm and the Julia 1.10 samplers:
//   rand(UInt64)          xoshiro256++ next()
//   rand(Float64)         Float64(u >>> 11) * 2^-53
//   rand(Float32)         Float32((u >>> 32) >>> 8) * 2^-24      (top 24 bits)
//   rand(rng, Float32, n) n*4 < 64 B -> xoshiro_bulk_nosimd: one u64 per 8 bytes,
//                         low 32 bits -> first float, high 32 bits -> second float,
//                         each Float32(u32 >>> 8) * 2^-24
//   rand(rng, Float64, n) n*8 < 64 B -> one u64 per element
//   rand(rng, Base.OneTo(n)::Int64)  SamplerRangeNDL (Lemire nearly-divisionless, UInt64)
// Call sites: CartPoleEnv.jl:99,101; PendulumEnv.jl:85-86; MountainCarEnv.jl:100;
// random_policy.jl:27,31; networks.jl:428 (Float64 uniforms for Gumbel-max).
#pragma once
#include <cstdint>

namespace jl {

struct Xoshiro {
    uint64_t s0, s1, s2, s3;
};

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static inline uint64_t next_u64(Xoshiro& g) {
    uint64_t s0 = g.s0, s1 = g.s1, s2 = g.s2, s3 = g.s3;
    uint64_t res = rotl64(s0 + s3, 23) + s0;
    uint64_t t = s1 << 17;
    s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t;
    s3 = rotl64(s3, 45);
    g.s0 = s0; g.s1 = s1; g.s2 = s2; g.s3 = s3;
    return res;
}

static inline double rand_f64(Xoshiro& g) { return (double)(next_u64(g) >> 11) * 0x1p-53; }
static inline float rand_f32(Xoshiro& g) {
    return (float)((uint32_t)(next_u64(g) >> 32) >> 8) * 0x1p-24f;
}
template <class T> static inline T rand_scalar(Xoshiro& g);
template <> inline float rand_scalar<float>(Xoshiro& g) { return rand_f32(g); }
template <> inline double rand_scalar<double>(Xoshiro& g) { return rand_f64(g); }

// rand(rng, T, 4) — the array API with 16/32 bytes (< 64 B SIMD threshold).
static inline void rand_array4(Xoshiro& g, float out[4]) {
    for (int k = 0; k < 2; ++k) {
        uint64_t u = next_u64(g);
        out[2 * k] = (float)((uint32_t)u >> 8) * 0x1p-24f;
        out[2 * k + 1] = (float)((uint32_t)(u >> 32) >> 8) * 0x1p-24f;
    }
}
static inline void rand_array4(Xoshiro& g, double out[4]) {
    for (int k = 0; k < 4; ++k) out[k] = rand_f64(g);
}

// rand(rng, Base.OneTo(n)) for Int64 n >= 1 -> value in 1..n
static inline int64_t rand_oneto(Xoshiro& g, uint64_t n) {
    unsigned __int128 m = (unsigned __int128)next_u64(g) * n;
    uint64_t l = (uint64_t)m;
    if (l < n) {
        uint64_t t = (0 - n) % n;
        while (l < t) {
            m = (unsigned __int128)next_u64(g) * n;
            l = (uint64_t)m;
        }
    }
    return (int64_t)(uint64_t)(m >> 64) + 1;
}

// Test-harness seeding (NOT Julia's Xoshiro(seed), which is SHA-based and version
// dependent): four successive splitmix64 outputs.  The C ABI takes raw 4xUInt64 states.
static inline uint64_t splitmix64(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline Xoshiro seed_splitmix(uint64_t seed) {
    Xoshiro g;
    g.s0 = splitmix64(seed); g.s1 = splitmix64(seed); g.s2 = splitmix64(seed); g.s3 = splitmix64(seed);
    return g;
}

}  // namespace jl
