// ORACLE — TEST INFRASTRUCTURE ONLY (see jl_math.hpp header).
//
// CPU restatement of the trajectory container + samplers on the hot path.
// The upstream code lives in the registered package ReinforcementLearningTrajectories
// (compat "0.4", RLCore/Project.toml:20,40) + CircularArrayBuffers ("0.1.12"), which is NOT
// vendored under /root/reference and has no Manifest pin -> PARITY UNPINNED.  What is
// restated is the published behaviour recorded in SURVEY Appendix B, anchored on the
// reference's own call sites and boundary tests:
//   push order / trace layout   RLCore/src/policies/agent/agent_base.jl:45-59,
//                               agent_srt_cache.jl:30-50, docs/src/How_to_implement_a_new_algorithm.md:84-112
//   length semantics            RLCore/test/policies/agent.jl:27-34 (0 after the first state, 1 after the first transition)
//   iteration tuple             RLCore/test/policies/q_based_policy.jl:40-58 (state,next_state,action,reward,terminal)
// Layout: a ring of cap+1 frames; a frame holds all `lanes` sub-envs (lanes = 1 is exactly
// the reference's single-stream CircularArraySARTSTraces).  Transition j uses state frame j
// and, as :next_state, state frame j+1 (MultiplexTraces).  Samplers draw WITH replacement
// (BatchSampler) from one Xoshiro stream per batch slot (a B200-side definition: the
// reference draws the whole batch from a single stream).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "jl_rng.hpp"

namespace oracle {

struct SumTree {  // leaves [L, 2L), node k = node 2k + node 2k+1, float32 sums
    int64_t L = 1;
    std::vector<float> tree;
    void init(int64_t n_slots) {
        L = 1;
        while (L < n_slots) L <<= 1;
        tree.assign(2 * L, 0.f);
    }
    float total() const { return tree[1]; }
    void set(int64_t slot, float p) {
        int64_t k = L + slot;
        tree[k] = p;
        for (k >>= 1; k >= 1; k >>= 1) tree[k] = tree[2 * k] + tree[2 * k + 1];
    }
    int64_t find(float v) const {
        int64_t k = 1;
        while (k < L) {
            int64_t l = 2 * k;
            if (v <= tree[l]) k = l;
            else { v -= tree[l]; k = l + 1; }
        }
        return k - L;
    }
};

struct Traj {
    int ns;
    int64_t lanes, cap;       // cap transitions frames; cap+1 physical frames
    int64_t first = 0, n_states = 0;
    std::vector<float> state;     // (ns, lanes, cap+1)
    std::vector<int32_t> action;  // (lanes, cap+1)
    std::vector<float> reward;    // (lanes, cap+1)
    std::vector<uint8_t> terminal;
    bool prioritized = false;
    float default_priority = 1.f;
    SumTree st;

    Traj(int ns_, int64_t lanes_, int64_t cap_, bool prio, float defp)
        : ns(ns_), lanes(lanes_), cap(cap_), state((size_t)ns_ * lanes_ * (cap_ + 1)), action(lanes_ * (cap_ + 1)),
          reward(lanes_ * (cap_ + 1)), terminal(lanes_ * (cap_ + 1)), prioritized(prio), default_priority(defp) {
        if (prio) st.init(lanes * (cap + 1));
    }
    int64_t frames() const { return cap + 1; }
    int64_t length() const { return n_states > 0 ? n_states - 1 : 0; }  // transition frames
    int64_t phys(int64_t j) const { return (first + j) % frames(); }
    // push!(trajectory, (state = s0,)) — agent_base.jl:45-47
    void push_state(const float* obs) {
        int64_t pf;
        if (n_states == frames()) {  // full: overwrite the oldest frame
            pf = first;
            first = (first + 1) % frames();
        } else {
            pf = phys(n_states);
            n_states += 1;
        }
        std::copy(obs, obs + (size_t)ns * lanes, state.begin() + (size_t)ns * lanes * pf);
        if (prioritized)
            for (int64_t e = 0; e < lanes; ++e) st.set(pf * lanes + e, 0.f);  // newest state: no transition yet
    }
    // push!(trajectory, (state = s', action, reward, terminal)) — agent_base.jl:56-59
    void push(const int32_t* a, const float* r, const uint8_t* t, const float* next_obs) {
        int64_t pf = phys(n_states - 1);  // frame of the state the action was taken in
        for (int64_t e = 0; e < lanes; ++e) {
            action[pf * lanes + e] = a[e];
            reward[pf * lanes + e] = r[e];
            terminal[pf * lanes + e] = t[e];
        }
        push_state(next_obs);
        if (prioritized)
            for (int64_t e = 0; e < lanes; ++e) st.set(pf * lanes + e, default_priority);
    }
    // gather one transition by logical flat index q = j*lanes + e
    void gather(int64_t q, float* s, int32_t* a, float* r, uint8_t* t, float* s2, int64_t* key) const {
        int64_t j = q / lanes, e = q % lanes;
        int64_t pf = phys(j), pn = phys(j + 1);
        for (int k = 0; k < ns; ++k) {
            s[k] = state[(size_t)ns * (pf * lanes + e) + k];
            s2[k] = state[(size_t)ns * (pn * lanes + e) + k];
        }
        *a = action[pf * lanes + e];
        *r = reward[pf * lanes + e];
        *t = terminal[pf * lanes + e];
        *key = pf * lanes + e;
    }
    int64_t logical_of_key(int64_t key) const {
        int64_t pf = key / lanes, e = key % lanes;
        int64_t j = (pf - first + frames()) % frames();
        return j * lanes + e;
    }
};

// BatchSampler: slot k draws rand(rng_k, 1:length) (with replacement)
static inline void sample_uniform(const Traj& tr, jl::Xoshiro* slots, int64_t B, int64_t* q_out) {
    uint64_t n = (uint64_t)(tr.length() * tr.lanes);
    for (int64_t k = 0; k < B; ++k) q_out[k] = jl::rand_oneto(slots[k], n) - 1;
}
// prioritised: slot k draws v = rand(rng_k, Float32) * total and descends the sum tree;
// weights w = (n * p / total)^(-beta) / max_k w   (PrioritizedDQN)
static inline void sample_prioritized(const Traj& tr, jl::Xoshiro* slots, int64_t B, float beta, int64_t* q_out, int64_t* key_out,
                                      float* prio_out, float* w_out) {
    float total = tr.st.total();
    int64_t n = tr.length() * tr.lanes;
    float wmax = 0.f;
    for (int64_t k = 0; k < B; ++k) {
        float v = jl::rand_f32(slots[k]) * total;
        int64_t key = tr.st.find(v);
        float p = tr.st.tree[tr.st.L + key];
        if (!(p > 0.f)) {  // rounding landed on an empty leaf: fall back to the oldest transition
            key = tr.phys(0) * tr.lanes;
            p = tr.st.tree[tr.st.L + key];
        }
        key_out[k] = key;
        prio_out[k] = p;
        q_out[k] = tr.logical_of_key(key);
        float w = std::pow((float)n * (p / total), -beta);
        w_out[k] = w;
        wmax = std::max(wmax, w);
    }
    for (int64_t k = 0; k < B; ++k) w_out[k] = w_out[k] / wmax;
}

}  // namespace oracle
