// ORACLE — TEST INFRASTRUCTURE ONLY (see jl_math.hpp header).
//
// CPU restatement of the reference's explorers for the DQN action path:
//   get_ϵ (linear | exp)                 RLCore/src/policies/explorers/epsilon_greedy_explorer.jl:69-91
//   plan!(::EpsilonGreedyExplorer, v)    :102-112  (is_break_tie true | false)
//   prob(::EpsilonGreedyExplorer, v)     :141-171
//   GreedyExplorer                       :196-204 (`findmax(values)[2]`)
//   BatchExplorer                        explorers/batch_explorer.jl:15-21 — the inner explorer is applied to
//                                        each column in turn, so column i sees step + i.
// Pinned by the reference's own vectors: RLCore/test/policies/explorers/epsilon_greedy_explorer.jl:8-74
// (schedule values, prob vectors, GreedyExplorer plan).  The stream arithmetic (rand(rng), rand(rng, 1:n),
// rand(rng, inds)) is Julia stdlib: PARITY UNPINNED (see jl_rng.hpp).
#pragma once
#include <cmath>
#include <cstdint>

#include "jl_rng.hpp"

namespace oracle {

struct Explorer {
    double eps_stable, eps_init;
    int64_t warmup_steps, decay_steps;
    int kind;           // 0 :linear, 1 :exp
    int is_break_tie;
};

// epsilon_greedy_explorer.jl:69-91.  All Float64, evaluated left to right, no contraction (-ffp-contract=off).
inline double get_eps(const Explorer& s, int64_t step) {
    if (s.kind == 0) {
        if (step <= s.warmup_steps) return s.eps_init;
        if (step >= s.warmup_steps + s.decay_steps) return s.eps_stable;
        int64_t steps_left = s.warmup_steps + s.decay_steps - step;
        return s.eps_stable + (double)steps_left / (double)s.decay_steps * (s.eps_init - s.eps_stable);
    }
    if (step <= s.warmup_steps) return s.eps_init;
    int64_t n = step - s.warmup_steps;
    double scale = s.eps_init - s.eps_stable;
    return s.eps_stable + scale * std::exp(-1.0 * (double)n / (double)s.decay_steps);
}

// findmax(values)[2] (0-based here): first maximum; NaN compares greater than everything (Base.isless order)
inline int findmax_index(const float* v, int n) {
    int best = 0;
    for (int o = 1; o < n; ++o) {
        bool gt = (std::isnan(v[o]) && !std::isnan(v[best])) || v[o] > v[best];
        if (gt) best = o;
    }
    return best;
}

// plan!(s::EpsilonGreedyExplorer, values) for one column (1-based action).  `eps` = get_ϵ(s) of this call.
// :105  rand(s.rng) >= ϵ ? rand(s.rng, find_all_max(values)[2]) : rand(s.rng, 1:length(values))
// :111  rand(s.rng) >= ϵ ? findmax(values)[2]                   : rand(s.rng, 1:length(values))
inline int egreedy_plan(const Explorer& s, double eps, const float* v, int n, jl::Xoshiro& g) {
    double u = jl::rand_f64(g);
    if (u >= eps) {
        if (!s.is_break_tie) return findmax_index(v, n) + 1;
        float mx = v[0];
        for (int o = 1; o < n; ++o) if (v[o] > mx) mx = v[o];
        int cnt = 0;
        for (int o = 0; o < n; ++o) cnt += v[o] == mx;
        int pick = (int)jl::rand_oneto(g, (uint64_t)cnt);   // rand(rng, inds) = inds[rand(rng, 1:length(inds))]
        for (int o = 0; o < n; ++o) {
            if (v[o] == mx && --pick == 0) return o + 1;
        }
        return n;   // unreachable for NaN-free input
    }
    return (int)jl::rand_oneto(g, (uint64_t)n);
}

// prob(s, values) -> probs[n]   (:141-171)
inline void egreedy_prob(const Explorer& s, double eps, const double* v, int n, double* probs) {
    for (int o = 0; o < n; ++o) probs[o] = eps / n;
    if (s.is_break_tie) {
        double mx = v[0];
        for (int o = 1; o < n; ++o) if (v[o] > mx) mx = v[o];
        int cnt = 0;
        for (int o = 0; o < n; ++o) cnt += v[o] == mx;
        for (int o = 0; o < n; ++o) if (v[o] == mx) probs[o] += (1 - eps) / cnt;
    } else {
        int best = 0;
        for (int o = 1; o < n; ++o) if (v[o] > v[best]) best = o;
        probs[best] += 1 - eps;
    }
}

}  // namespace oracle
