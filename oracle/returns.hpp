// ORACLE — TEST INFRASTRUCTURE ONLY (see jl_math.hpp header).
//
// Restatement of /root/reference/src/ReinforcementLearningCore/src/utils/basic.jl:
//   discount_rewards(!)            :138-235  (core loop :227-235)
//   discount_rewards_reduced(!)    :237-319
//   generalized_advantage_estimation(!) :334-417 (core loop :408-417)
// PINNED by the reference's own golden vectors
// (src/ReinforcementLearningCore/test/utils/base.jl:22-152) in tests/test_oracle_returns.py.
//
// Matrices are Julia column-major: element (i, j) of an (R, C) matrix is at i + R*j.
// `dims = k` means TIME RUNS ALONG DIM k (basic.jl flips it for eachslice, :152,:386):
//   dims = 1: each column is a series   (series s, time i) at  i + R*s
//   dims = 2: each row is a series      (series s, time i) at  s + R*i
// `values` has one more entry than `rewards` along the time dim.
// Operation order is exactly the reference's:
//   gain  = r[i] + (gamma * gain) * c
//   delta = (r[i] + (gamma * v[i+1]) * c) - v[i];  gae = delta + ((gamma*lambda) * c) * gae
// with c::Bool multiplication being Julia's strong zero.  -ffp-contract=off.
#pragma once
#include <cstdint>

#include "jl_math.hpp"

namespace oracle {

// strided vector kernels ------------------------------------------------------------
template <class T>
void discount_rewards_vec(T* out, int64_t so, const T* r, int64_t sr, const uint8_t* term,
                          int64_t st, T gamma, T init, int64_t n) {
    T gain = init;
    for (int64_t i = n - 1; i >= 0; --i) {
        bool cont = term ? !term[i * st] : true;
        gain = r[i * sr] + jl::mul_bool(gamma * gain, cont);
        out[i * so] = gain;
    }
}
template <class T>
T discount_rewards_reduced_vec(const T* r, int64_t sr, const uint8_t* term, int64_t st, T gamma,
                               T init, int64_t n) {
    T gain = init;
    for (int64_t i = n - 1; i >= 0; --i) {
        bool cont = term ? !term[i * st] : true;
        gain = r[i * sr] + jl::mul_bool(gamma * gain, cont);
    }
    return gain;
}
template <class T>
void gae_vec(T* adv, int64_t sa, const T* r, int64_t sr, const T* v, int64_t sv,
             const uint8_t* term, int64_t st, T gamma, T lambda, int64_t n) {
    T gae = 0;
    for (int64_t i = n - 1; i >= 0; --i) {
        bool cont = term ? !term[i * st] : true;
        T delta = (r[i * sr] + jl::mul_bool(gamma * v[(i + 1) * sv], cont)) - v[i * sv];
        gae = delta + jl::mul_bool(gamma * lambda, cont) * gae;
        adv[i * sa] = gae;
    }
}

// matrix front-ends (rewards is (R, C) column-major) ----------------------------------
template <class T>
void discount_rewards_mat(T* out, const T* r, const uint8_t* term, const T* init, T gamma,
                          int64_t R, int64_t C, int dims) {
    int64_t n_series = dims == 1 ? C : R, n_time = dims == 1 ? R : C;
    int64_t s_series = dims == 1 ? R : 1, s_time = dims == 1 ? 1 : R;
    for (int64_t s = 0; s < n_series; ++s)
        discount_rewards_vec<T>(out + s * s_series, s_time, r + s * s_series, s_time,
                                term ? term + s * s_series : nullptr, s_time, gamma,
                                init ? init[s] : (T)0, n_time);
}
template <class T>
void discount_rewards_reduced_mat(T* out, const T* r, const uint8_t* term, const T* init, T gamma,
                                  int64_t R, int64_t C, int dims) {
    int64_t n_series = dims == 1 ? C : R, n_time = dims == 1 ? R : C;
    int64_t s_series = dims == 1 ? R : 1, s_time = dims == 1 ? 1 : R;
    for (int64_t s = 0; s < n_series; ++s)
        out[s] = discount_rewards_reduced_vec<T>(r + s * s_series, s_time,
                                                 term ? term + s * s_series : nullptr, s_time,
                                                 gamma, init ? init[s] : (T)0, n_time);
}
// values is (R+1, C) for dims = 1 and (R, C+1) for dims = 2.
template <class T>
void gae_mat(T* adv, const T* r, const T* v, const uint8_t* term, T gamma, T lambda, int64_t R,
             int64_t C, int dims) {
    if (dims == 1) {
        for (int64_t s = 0; s < C; ++s)
            gae_vec<T>(adv + s * R, 1, r + s * R, 1, v + s * (R + 1), 1,
                       term ? term + s * R : nullptr, 1, gamma, lambda, R);
    } else {
        for (int64_t s = 0; s < R; ++s)
            gae_vec<T>(adv + s, R, r + s, R, v + s, R, term ? term + s : nullptr, R, gamma,
                       lambda, C);
    }
}

}  // namespace oracle
