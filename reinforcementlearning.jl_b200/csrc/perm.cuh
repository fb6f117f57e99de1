// perm.cuh — minibatch permutation generated on the device (used when the host does not pass
// its own `shuffle!(rng, 1:N*T)` result, SURVEY Appendix B PPO `_update!`).
//
// A keyed 4-round alternating (unbalanced) Feistel network on the smallest power-of-two
// domain 2^bits >= n, bits >= 2, with cycle walking for the values >= n.  The left half has
// floor(bits/2) bits and the right half the rest; the halves swap widths every round, so any
// bits works and a power-of-two n (the BASELINE rollouts: 65536 x 32 = 2^21) never walks.
// DESIGN.md §K7; the parity tests check it against an independent CPU restatement.
#pragma once
#include <cstdint>

namespace b200perm {

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// smallest bits >= 2 with 2^bits >= n   (n >= 1)
__host__ __device__ __forceinline__ int perm_bits(uint32_t n) {
    int bits = 2;
    while (bits < 32 && (1u << bits) < n) ++bits;
    return bits;
}

__host__ __device__ __forceinline__ uint32_t perm_index_bits(uint32_t q, uint32_t n, uint32_t key, int bits) {
    const int a = bits >> 1, b = bits - a;
    const uint32_t mask_a = (1u << a) - 1u, mask_b = (1u << b) - 1u;
    uint32_t x = q;
    do {
        uint32_t l = x >> b, r = x & mask_b;
        uint32_t t;
        t = l ^ (mix32(r + key) & mask_a); l = r; r = t;                     // l: b bits, r: a bits
        t = l ^ (mix32(r + key + 0x9E3779B9u) & mask_b); l = r; r = t;       // l: a bits, r: b bits
        t = l ^ (mix32(r + key + 2u * 0x9E3779B9u) & mask_a); l = r; r = t;
        t = l ^ (mix32(r + key + 3u * 0x9E3779B9u) & mask_b); l = r; r = t;
        x = (l << b) | r;
    } while (x >= n);
    return x;
}

__host__ __device__ __forceinline__ uint32_t perm_index(uint32_t q, uint32_t n, uint32_t key) {
    return perm_index_bits(q, n, key, perm_bits(n));
}

}  // namespace b200perm
