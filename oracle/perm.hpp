// ORACLE — TEST INFRASTRUCTURE ONLY (see jl_math.hpp header).
// Minibatch permutation used when the host does not supply `shuffle!(rng, 1:N*T)` itself
// (PPO _update!): a keyed 4-round Feistel bijection on the next even
// power-of-two domain with cycle walking.  A B200-side definition (the reference's shuffle! is
// a sequential Fisher–Yates on one stream); DESIGN.md §K7.
#pragma once
#include <cstdint>
namespace oracle {
static inline uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
static inline uint32_t perm_index(uint32_t q, uint32_t n, uint32_t key) {
    int bits = 2;
    while ((1ull << bits) < n) bits += 2;
    int hb = bits / 2;
    uint32_t mask = (1u << hb) - 1;
    uint32_t x = q;
    do {
        uint32_t l = x >> hb, r = x & mask;
        for (uint32_t round = 0; round < 4; ++round) {
            uint32_t t = l ^ (mix32(r + key + round * 0x9E3779B9u) & mask);
            l = r; r = t;
        }
        x = (l << hb) | r;
    } while (x >= n);
    return x;
}
}  // namespace oracle
