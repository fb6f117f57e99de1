// ORACLE — TEST INFRASTRUCTURE ONLY (see jl_math.hpp header).
//
// Vector env in the shape of the reference's historical `MultiThreadEnv` (absent from the
// snapshot; described in docs/homepage/blog/an_introduction_to_reinforcement_learning_jl_
// design_implementations_thoughts/index.md:347-378 and SURVEY Appendix B): an array of
// heap-allocated per-env objects, a parallel-for over envs per step (OpenMP static
// schedule standing in for Threads.@threads), results gathered into batch arrays, and a
// soft reset of finished sub-envs.  PARITY UNPINNED (component absent upstream).
//
// Flag byte per env: bit0 = terminal after the last act!, bit1 = already re-initialised by
// the fused auto-reset (so a later soft reset must not draw from the RNG again).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "envs.hpp"

namespace oracle {

enum Field { F_STATE = 0, F_OBS = 1, F_REWARD = 2, F_TERMINAL = 3, F_T = 4, F_RNG = 5, F_FLAGS = 6, F_ACTION = 7 };

struct VecEnvBase {
    virtual ~VecEnvBase() {}
    virtual void reset(int force) = 0;
    virtual int step(const void* actions, int auto_reset) = 0;                  // returns #invalid actions
    virtual int step_random(int auto_reset, int32_t* actions_out) = 0;          // RandomPolicy on the env stream
    virtual void get(int field, void* dst) const = 0;
    virtual void set(int field, const void* src) = 0;
    virtual int64_t size() const = 0;
    // MaxTimeoutEnv(env, max_t) (wrappers/MaxTimeoutEnv.jl:17-28): is_terminated also when
    // current_t > max_t; current_t starts at 1 with every reset! and counts act! calls, so
    // current_t == env.t + 1.  reward(env) still forwards to the wrapped env.  0 = no wrapper.
    int64_t max_timeout = 0;
};

template <class E, class T, class ActT> struct VecEnv : VecEnvBase {
    std::vector<std::unique_ptr<E>> envs;
    std::vector<uint8_t> flags;
    std::vector<T> rewards;
    template <class P> VecEnv(int64_t N, const P& p, const uint64_t* rng) : flags(N, 0), rewards(N, (T)0) {
        envs.reserve(N);
        for (int64_t i = 0; i < N; ++i) {
            jl::Xoshiro g{rng[4 * i], rng[4 * i + 1], rng[4 * i + 2], rng[4 * i + 3]};
            envs.emplace_back(new E(p, g, true));
            reset_reward(i);
        }
    }
    int64_t size() const override { return (int64_t)envs.size(); }
    void reset(int force) override {
        int64_t N = size();
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            if (force || ((flags[i] & 1) && !(flags[i] & 2))) { envs[i]->reset(); reset_reward(i); }
            flags[i] = 0;
        }
    }
    static bool do_act(E& e, ActT a);
    static int64_t n_random_actions(const E& e);
    void reset_reward(int64_t) {}   // specialised for envs whose reset! sets the reward field (AcrobotEnv.jl:105)
    void post(int64_t i, int auto_reset) {
        E& e = *envs[i];
        rewards[i] = e.reward();
        bool term = e.done || (max_timeout > 0 && e.t + 1 > max_timeout);
        uint8_t f = term ? 1 : 0;
        if (auto_reset && term) { e.reset(); f = 3; }
        flags[i] = f;
    }
    int step(const void* actions, int auto_reset) override {
        const ActT* a = (const ActT*)actions;
        int64_t N = size();
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (int64_t i = 0; i < N; ++i) {
            if (!do_act(*envs[i], a[i])) { bad += 1; continue; }
            post(i, auto_reset);
        }
        return bad;
    }
    int step_random(int auto_reset, int32_t* actions_out) override {
        int64_t N = size();
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            E& e = *envs[i];
            int64_t a = jl::rand_oneto(e.rng, (uint64_t)n_random_actions(e));  // random_policy.jl:29-32
            if (actions_out) actions_out[i] = (int32_t)a;
            do_act_discrete(e, a);
            post(i, auto_reset);
        }
        return 0;
    }
    static void do_act_discrete(E& e, int64_t a);
    void get(int field, void* dst) const override {
        int64_t N = size();
        switch (field) {
            case F_STATE: for (int64_t i = 0; i < N; ++i) for (int k = 0; k < E::NS; ++k) ((T*)dst)[i * E::NS + k] = envs[i]->state[k]; break;
            case F_OBS: for (int64_t i = 0; i < N; ++i) envs[i]->obs((T*)dst + i * E::NOBS); break;
            case F_REWARD: for (int64_t i = 0; i < N; ++i) ((T*)dst)[i] = rewards[i]; break;
            case F_TERMINAL: for (int64_t i = 0; i < N; ++i) ((uint8_t*)dst)[i] = flags[i] & 1; break;
            case F_FLAGS: for (int64_t i = 0; i < N; ++i) ((uint8_t*)dst)[i] = flags[i]; break;
            case F_T: for (int64_t i = 0; i < N; ++i) ((int32_t*)dst)[i] = (int32_t)envs[i]->t; break;
            case F_RNG: for (int64_t i = 0; i < N; ++i) { const jl::Xoshiro& g = envs[i]->rng; uint64_t* d = (uint64_t*)dst + 4 * i; d[0] = g.s0; d[1] = g.s1; d[2] = g.s2; d[3] = g.s3; } break;
            case F_ACTION: for (int64_t i = 0; i < N; ++i) action_out(*envs[i], dst, i); break;   // env.action (reset! redraws it: CartPoleEnv.jl:101)
            default: break;
        }
    }
    static void action_out(const E& e, void* dst, int64_t i);
    void set(int field, const void* src) override {
        int64_t N = size();
        switch (field) {
            case F_STATE: for (int64_t i = 0; i < N; ++i) for (int k = 0; k < E::NS; ++k) envs[i]->state[k] = ((const T*)src)[i * E::NS + k]; break;
            case F_T: for (int64_t i = 0; i < N; ++i) envs[i]->t = ((const int32_t*)src)[i]; break;
            case F_RNG: for (int64_t i = 0; i < N; ++i) { const uint64_t* d = (const uint64_t*)src + 4 * i; envs[i]->rng = jl::Xoshiro{d[0], d[1], d[2], d[3]}; } break;
            case F_FLAGS: for (int64_t i = 0; i < N; ++i) { flags[i] = ((const uint8_t*)src)[i]; envs[i]->done = flags[i] & 1; } break;
            default: break;
        }
    }
};

using VecCartPoleF32 = VecEnv<CartPole<float>, float, int32_t>;
using VecCartPoleF64 = VecEnv<CartPole<double>, double, int32_t>;
using VecPendulumC = VecEnv<Pendulum, float, float>;     // continuous torque
using VecPendulumD = VecEnv<Pendulum, float, int32_t>;   // discrete torque index
using VecMountainCar = VecEnv<MountainCar, float, int32_t>;
using VecCartPoleC = VecEnv<CartPole<float, true>, float, float>;   // CartPoleEnv(continuous = true)
using VecMountainCarC = VecEnv<MountainCar, float, float>;          // ContinuousMountainCarEnv
// T = Float64 (the reference constructors' default for Pendulum / MountainCar); a continuous action is a Float64 then
struct Pendulum64C : PendulumT<double> { using PendulumT<double>::PendulumT; };
struct Pendulum64D : PendulumT<double> { using PendulumT<double>::PendulumT; };
struct MountainCar64 : MountainCarT<double> { using MountainCarT<double>::MountainCarT; };
struct MountainCar64C : MountainCarT<double> { using MountainCarT<double>::MountainCarT; };
using VecPendulum64C = VecEnv<Pendulum64C, double, double>;
using VecPendulum64D = VecEnv<Pendulum64D, double, int32_t>;
using VecMountainCar64 = VecEnv<MountainCar64, double, int32_t>;
using VecMountainCar64C = VecEnv<MountainCar64C, double, double>;
using VecAcrobot = VecEnv<Acrobot, double, int32_t>;   // AcrobotEnv{Float64}
template <> inline void VecAcrobot::reset_reward(int64_t i) { rewards[i] = envs[i]->reward(); }
template <> inline void VecAcrobot::action_out(const Acrobot& e, void* d, int64_t i) { ((int32_t*)d)[i] = (int32_t)e.action; }
template <> inline bool VecAcrobot::do_act(Acrobot& e, int32_t a) { return e.act(a); }
template <> inline int64_t VecAcrobot::n_random_actions(const Acrobot&) { return 3; }
template <> inline void VecAcrobot::do_act_discrete(Acrobot& e, int64_t a) { e.act(a); }
template <> inline void VecPendulum64C::action_out(const Pendulum64C& e, void* d, int64_t i) { ((double*)d)[i] = e.action; }
template <> inline void VecPendulum64D::action_out(const Pendulum64D& e, void* d, int64_t i) { ((float*)d)[i] = (float)e.action; }
template <> inline void VecMountainCar64::action_out(const MountainCar64& e, void* d, int64_t i) { ((int32_t*)d)[i] = (int32_t)e.action; }
template <> inline void VecMountainCar64C::action_out(const MountainCar64C& e, void* d, int64_t i) { ((double*)d)[i] = e.action_f; }
template <> inline bool VecPendulum64C::do_act(Pendulum64C& e, double a) { return e.act_continuous(a); }
template <> inline bool VecPendulum64D::do_act(Pendulum64D& e, int32_t a) { return e.act_discrete(a); }
template <> inline bool VecMountainCar64::do_act(MountainCar64& e, int32_t a) { return e.act(a); }
template <> inline bool VecMountainCar64C::do_act(MountainCar64C& e, double a) { return e.act_continuous(a); }
template <> inline int64_t VecPendulum64C::n_random_actions(const Pendulum64C& e) { return e.n_actions; }
template <> inline int64_t VecPendulum64D::n_random_actions(const Pendulum64D& e) { return e.n_actions; }
template <> inline int64_t VecMountainCar64::n_random_actions(const MountainCar64&) { return 3; }
template <> inline int64_t VecMountainCar64C::n_random_actions(const MountainCar64C&) { return 0; }
template <> inline void VecPendulum64C::do_act_discrete(Pendulum64C& e, int64_t a) { e.act_discrete(a); }
template <> inline void VecPendulum64D::do_act_discrete(Pendulum64D& e, int64_t a) { e.act_discrete(a); }
template <> inline void VecMountainCar64::do_act_discrete(MountainCar64& e, int64_t a) { e.act(a); }
template <> inline void VecMountainCar64C::do_act_discrete(MountainCar64C&, int64_t) {}

template <> inline void VecCartPoleF32::action_out(const CartPole<float>& e, void* d, int64_t i) { ((int32_t*)d)[i] = (int32_t)e.action; }
template <> inline void VecCartPoleF64::action_out(const CartPole<double>& e, void* d, int64_t i) { ((int32_t*)d)[i] = (int32_t)e.action; }
template <> inline void VecMountainCar::action_out(const MountainCar& e, void* d, int64_t i) { ((int32_t*)d)[i] = (int32_t)e.action; }
template <> inline void VecPendulumC::action_out(const Pendulum& e, void* d, int64_t i) { ((float*)d)[i] = (float)e.action; }
template <> inline void VecPendulumD::action_out(const Pendulum& e, void* d, int64_t i) { ((float*)d)[i] = (float)e.action; }
template <> inline void VecCartPoleC::action_out(const CartPole<float, true>& e, void* d, int64_t i) { ((float*)d)[i] = (float)e.action_f; }
template <> inline void VecMountainCarC::action_out(const MountainCar& e, void* d, int64_t i) { ((float*)d)[i] = (float)e.action_f; }
template <> inline bool VecCartPoleF32::do_act(CartPole<float>& e, int32_t a) { return e.act(a); }
template <> inline bool VecCartPoleF64::do_act(CartPole<double>& e, int32_t a) { return e.act(a); }
template <> inline bool VecPendulumC::do_act(Pendulum& e, float a) { return e.act_continuous((double)a); }
template <> inline bool VecPendulumD::do_act(Pendulum& e, int32_t a) { return e.act_discrete(a); }
template <> inline bool VecMountainCar::do_act(MountainCar& e, int32_t a) { return e.act(a); }
template <> inline bool VecCartPoleC::do_act(CartPole<float, true>& e, float a) { return e.act_continuous(a); }
template <> inline bool VecMountainCarC::do_act(MountainCar& e, float a) { return e.act_continuous(a); }
template <> inline int64_t VecCartPoleC::n_random_actions(const CartPole<float, true>&) { return 0; }
template <> inline int64_t VecMountainCarC::n_random_actions(const MountainCar&) { return 0; }
template <> inline void VecCartPoleC::do_act_discrete(CartPole<float, true>&, int64_t) {}
template <> inline void VecMountainCarC::do_act_discrete(MountainCar&, int64_t) {}
template <> inline int64_t VecCartPoleF32::n_random_actions(const CartPole<float>&) { return 2; }
template <> inline int64_t VecCartPoleF64::n_random_actions(const CartPole<double>&) { return 2; }
template <> inline int64_t VecPendulumC::n_random_actions(const Pendulum& e) { return e.n_actions; }
template <> inline int64_t VecPendulumD::n_random_actions(const Pendulum& e) { return e.n_actions; }
template <> inline int64_t VecMountainCar::n_random_actions(const MountainCar&) { return 3; }
template <> inline void VecCartPoleF32::do_act_discrete(CartPole<float>& e, int64_t a) { e.act(a); }
template <> inline void VecCartPoleF64::do_act_discrete(CartPole<double>& e, int64_t a) { e.act(a); }
template <> inline void VecPendulumC::do_act_discrete(Pendulum& e, int64_t a) { e.act_discrete(a); }
template <> inline void VecPendulumD::do_act_discrete(Pendulum& e, int64_t a) { e.act_discrete(a); }
template <> inline void VecMountainCar::do_act_discrete(MountainCar& e, int64_t a) { e.act(a); }

}  // namespace oracle
