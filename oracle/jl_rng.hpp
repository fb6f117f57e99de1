// ORACLE — TEST INFRASTRUCTURE ONLY (see jl_math.hpp header).
//
// CPU restatement of the Julia stdlib `Random` pieces the reference's envs/policies
// consume (Julia 1.10: stdlib/Random/src/Xoshiro.jl, XoshiroSimd.jl, generation.jl).
// `Random` is NOT under /root/reference (Julia stdlib, version pinned only by
// .devcontainer/Dockerfile:1 `julia:1.10`); restated from the published xoshiro256++ algorithm and the Julia 1.10
// samplers.  PINNED by the known answers Julia's own manual prints (stdlib Random docstrings; tests/test_oracle_julia_rng.py):
//   Xoshiro(1234); rand(rng, 2) == [0.32597672886359486, 0.5490511363155669]     (generator core, integer seeding, Float64 sampler)
//   shuffle(Xoshiro(123), Vector(1:10)) == [5,4,2,3,6,10,8,1,9,7]; randperm(Xoshiro(123), 4) == [1,4,2,3];
//   randcycle(Xoshiro(123), 6) == [5,4,2,6,3,1]                                   (UInt52Raw = u64 >>> 12, masked rejection)
// Still unpinned: the Float32 samplers, the array path below 64 bytes, rand(rng, Base.OneTo(n)) (Lemire), randn.
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

// rand(rng, UInt52Raw()) for Xoshiro = rand(rng, UInt64) >>> 12, and the masked rejection sampler ltm52(n, mask) =
// LessThan(n - 1, Masked(mask, UInt52Raw(Int))) that shuffle! / randperm! / randcycle! draw from (stdlib Random, misc.jl).
static inline uint64_t rand_ltm52(Xoshiro& g, uint64_t n, uint64_t mask) {
    for (;;) {
        uint64_t x = (next_u64(g) >> 12) & mask;
        if (x <= n - 1) return x;
    }
}
// shuffle!(rng, a) (stdlib Random misc.jl): for i = 2:n  j = 1 + rand(rng, ltm52(i, mask)); swap a[i], a[j]; mask grows with i.
// This is the `shuffle!(rng, 1:N*T)` of the PPO update — pinned by Julia's documented
// shuffle(Xoshiro(123), Vector(1:10)) == [5, 4, 2, 3, 6, 10, 8, 1, 9, 7] (tests/test_oracle_julia_rng.py).
template <class T> static inline void shuffle(Xoshiro& g, T* a, int64_t n) {
    uint64_t mask = 3;
    for (int64_t i = 2; i <= n; ++i) {
        int64_t j = 1 + (int64_t)rand_ltm52(g, (uint64_t)i, mask);
        T tmp = a[i - 1]; a[i - 1] = a[j - 1]; a[j - 1] = tmp;
        if ((uint64_t)i == 1 + mask) mask = 2 * mask + 1;
    }
}
// randperm!(rng, a)
static inline void randperm(Xoshiro& g, int64_t* a, int64_t n) {
    if (n == 0) return;
    a[0] = 1;
    uint64_t mask = 3;
    for (int64_t i = 2; i <= n; ++i) {
        int64_t j = 1 + (int64_t)rand_ltm52(g, (uint64_t)i, mask);
        if (i != j) a[i - 1] = a[j - 1];
        a[j - 1] = i;
        if ((uint64_t)i == 1 + mask) mask = 2 * mask + 1;
    }
}

// Test-harness seeding: four successive splitmix64 outputs.  (Julia's own Xoshiro(seed::Integer) for 1.7 <= version <= 1.10 is
// the SHA-256 digest of the seed's UInt32 words read as four little-endian UInt64 — restated with hashlib in
// tests/oracle_lib.julia_xoshiro and pinned by Julia's documented Xoshiro(1234) vector.)  The C ABI takes raw 4xUInt64 states.
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
