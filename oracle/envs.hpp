// ORACLE — TEST INFRASTRUCTURE ONLY (see jl_math.hpp header).
//
// Scalar, one-object-per-env restatement of the reference's classic-control envs, with
// Julia's promotion rules spelled out.  Follows
//   CartPoleEnv.jl:22-46 (params), :84-86 (reward/terminated/state), :98-104 (reset!),
//                  :112-116 (act!), :118-140 (_step!)
//   PendulumEnv.jl:41-66, :70-71, :80-92, :94-122
//   MountainCarEnv.jl:19-40, :95-105, :113-135
// under /root/reference/src/ReinforcementLearningEnvironments/src/environments/examples/.
// Compile with -ffp-contract=off: Julia never contracts a*b+c outside explicit muladd.
// PARITY UNPINNED by the reference (no golden trajectories exist); pinned by
// the source-derived known answers of SURVEY Appendix C (tests/test_oracle_envs.py).
#pragma once
#include <cstdint>
#include <vector>

#include "jl_math.hpp"
#include "jl_rng.hpp"

namespace oracle {

// Final field values of the reference's params structs (already converted to T by the
// reference constructor; carried as double, which embeds Float32 exactly).
struct CartPoleParams {  // CartPoleEnvParams{T}, CartPoleEnv.jl:3-15
    double gravity, masscart, masspole, totalmass, halflength, polemasslength, forcemag, dt,
        thetathreshold, xthreshold;
    int64_t max_steps;
};
template <class T> inline CartPoleParams cartpole_default_params() {
    // CartPoleEnvParams{T}(; ...) CartPoleEnv.jl:22-46: derived fields computed in Float64,
    // every field converted to T by the struct constructor.
    CartPoleParams p;
    p.gravity = (T)9.8; p.masscart = (T)1.0; p.masspole = (T)0.1;
    p.totalmass = (T)(1.0 + 0.1); p.halflength = (T)0.5; p.polemasslength = (T)(0.1 * 0.5);
    p.forcemag = (T)10.0; p.dt = (T)0.02; p.thetathreshold = (T)(12.0 * jl::PI_D / 180);
    p.xthreshold = (T)2.4; p.max_steps = 200;
    return p;
}

// CONT = true is CartPoleEnv(continuous = true): ACT = T, action_space -1.0..1.0
// (CartPoleEnv.jl:74-79,96,106-110).
template <class T, bool CONT = false> struct CartPole {
    static constexpr int NS = 4, NOBS = 4, NACT = 2;
    T action_f = 0;  // env.action for the continuous variant
    T g, M, m, l, pml, fmag, dt, ththr, xthr;
    int64_t max_steps;
    std::vector<T> state;  // heap vector per env, like the reference's Vector{T}
    int64_t action = 0;
    bool done = false;
    int64_t t = 0;
    jl::Xoshiro rng;

    CartPole(const CartPoleParams& p, jl::Xoshiro r, bool do_reset = true) : state(4, (T)0), rng(r) {
        g = (T)p.gravity; M = (T)p.totalmass; m = (T)p.masspole; l = (T)p.halflength;
        pml = (T)p.polemasslength; fmag = (T)p.forcemag; dt = (T)p.dt;
        ththr = (T)p.thetathreshold; xthr = (T)p.xthreshold; max_steps = p.max_steps;
        if (do_reset) reset();  // CartPoleEnv.jl:77
    }
    void reset() {  // CartPoleEnv.jl:98-104
        T u[4];
        jl::rand_array4(rng, u);  // rand(rng, T, 4)
        for (int k = 0; k < 4; ++k) state[k] = (T)0.1 * u[k] - (T)0.05;
        t = 0;
        if (CONT) {
            // rand(rng, -1.0..1.0): DomainSets' interval sampler (external, UNPINNED) — restated as
            // leftendpoint + rand(rng, Float64) * width, i.e. one Float64 draw = one 64-bit output.
            action_f = (T)(-1.0 + jl::rand_f64(rng) * 2.0);
        } else {
            action = jl::rand_oneto(rng, 2);  // rand(rng, Base.OneTo(2)) — consumes the stream
        }
        done = false;
    }
    bool act_continuous(T a) {  // CartPoleEnv.jl:106-110; `a in -1.0..1.0` (NaN fails)
        if (!(a >= (T)-1 && a <= (T)1)) return false;
        action_f = a;
        step_force(a * fmag);  // _step!(env, a): force = a * forcemag
        return true;
    }
    bool act(int64_t a) {  // CartPoleEnv.jl:112-116; returns false on `@assert` failure
        if (a < 1 || a > 2) return false;
        action = a;
        step(a == 2 ? 1 : -1);
        return true;
    }
    void step(int a) { step_force((T)a * fmag); }
    void step_force(T force) {  // CartPoleEnv.jl:118-140
        t += 1;
        T x = state[0], xdot = state[1], theta = state[2], thetadot = state[3];
        T c = jl::jcos(theta), s = jl::jsin(theta);
        T tmp = (force + (pml * (thetadot * thetadot)) * s) / M;
        // `4 / 3` is a Float64 literal expression: everything touching it is Float64.
        double den = (double)l * (4.0 / 3.0 - (double)((m * (c * c)) / M));
        double thetaacc = (double)(g * s - c * tmp) / den;
        double xacc = (double)tmp - (((double)pml * thetaacc) * (double)c) / (double)M;
        state[0] = x + dt * xdot;
        state[1] = (T)((double)xdot + (double)dt * xacc);
        state[2] = theta + dt * thetadot;
        state[3] = (T)((double)thetadot + (double)dt * thetaacc);
        done = std::fabs(state[0]) > xthr || std::fabs(state[2]) > ththr || t > max_steps;
    }
    T reward() const { return done ? (T)0 : (T)1; }  // CartPoleEnv.jl:84
    void obs(T* out) const { for (int k = 0; k < 4; ++k) out[k] = state[k]; }
};

struct PendulumParams {  // PendulumEnvParams{T}, PendulumEnv.jl:3-11 (+ n_actions, continuous)
    double max_speed, max_torque, g, m, l, dt;
    int64_t max_steps;
    int64_t n_actions;
    int32_t continuous;
};
inline PendulumParams pendulum_default_params() {
    return PendulumParams{8, 2, 10, 1, 1, (double)0.05f, 200, 3, 1};
}

// T = Float32 | Float64 (PendulumEnv.jl:42: the constructor's default is Float64).  Written so that the Float32 instantiation keeps
// Julia's promotion points (2*pi, mod() and the cost terms are Float64) and the Float64 one is all-double.
template <class T_> struct PendulumT {
    using T = T_;
    static constexpr int NS = 2, NOBS = 3;
    T max_speed, max_torque, g, m, l, dt;
    int64_t max_steps, n_actions;
    bool continuous;
    std::vector<T> state;
    T action = 0, rew = 0;
    bool done = false;
    int64_t t = 0;
    jl::Xoshiro rng;

    PendulumT(const PendulumParams& p, jl::Xoshiro r, bool do_reset = true) : state(2, (T)0), rng(r) {
        max_speed = (T)p.max_speed; max_torque = (T)p.max_torque; g = (T)p.g; m = (T)p.m;
        l = (T)p.l; dt = (T)p.dt; max_steps = p.max_steps; n_actions = p.n_actions;
        continuous = p.continuous != 0;
        if (do_reset) reset();  // PendulumEnv.jl:64
    }
    void reset() {  // PendulumEnv.jl:84-92; two scalar rand(rng, T) draws
        T u1 = jl::rand_scalar<T>(rng);
        state[0] = (T)((2 * jl::PI_D) * (double)(u1 - (T)1));  // 2*pi is Float64
        T u2 = jl::rand_scalar<T>(rng);
        state[1] = (T)2 * (u2 - (T)1);
        action = 0; t = 0; done = false; rew = 0;
    }
    // continuous: a in -2.0..2.0 (PendulumEnv.jl:73,95,122)
    bool act_continuous(double a) {
        if (!(a >= -2.0 && a <= 2.0)) return false;
        action = (T)a;
        step(action);
        return true;
    }
    // discrete: a in 1..n_actions, torque() PendulumEnv.jl:120-121 (Float64, stored as T)
    bool act_discrete(int64_t a) {
        if (a < 1 || a > n_actions) return false;
        double tq = (4.0 / (double)(n_actions - 1)) * ((double)a - (double)(n_actions - 1) / 2 - 1);
        action = (T)tq;
        step(action);
        return true;
    }
    void step(T a) {  // PendulumEnv.jl:100-118
        t += 1;
        T th = state[0], thdot = state[1];
        a = jl::jclamp(a, -max_torque, max_torque);
        // angle_normalize: T + pi -> T; mod(T, Float64) -> Float64
        double an = jl::jmod((double)(th + (T)jl::PI_D), 2 * jl::PI_D) - jl::PI_D;
        double costs = (an * an + 0.1 * (double)(thdot * thdot)) + 0.001 * (double)(a * a);
        T newthdot = thdot + (((((T)-3 * g) / ((T)2 * l)) * jl::jsin(th + (T)jl::PI_D)) +
                              (((T)3 * a) / (m * (l * l)))) * dt;
        th = th + newthdot * dt;
        newthdot = jl::jclamp(newthdot, -max_speed, max_speed);
        state[0] = th; state[1] = newthdot;
        done = t >= max_steps;
        rew = (T)(-costs);
    }
    T reward() const { return rew; }
    void obs(T* out) const {  // pendulum_observation PendulumEnv.jl:70
        out[0] = jl::jsin(state[0]); out[1] = jl::jcos(state[0]); out[2] = state[1];
    }
};
using Pendulum = PendulumT<float>;

struct MountainCarParams {  // MountainCarEnvParams{T}, MountainCarEnv.jl:3-12
    double min_pos, max_pos, max_speed, goal_pos, goal_velocity, power, gravity;
    int64_t max_steps;
};
inline MountainCarParams mountaincar_default_params() {
    using T = float;
    return MountainCarParams{(T)-1.2, (T)0.6, (T)0.07, (T)0.5, (T)0.0, (T)0.001, (T)0.0025, 200};
}

inline MountainCarParams mountaincar_continuous_default_params() {  // MountainCarEnv.jl:73-74
    using T = float;
    return MountainCarParams{(T)-1.2, (T)0.6, (T)0.07, (T)0.45, (T)0.0, (T)0.0015, (T)0.0025, 200};
}

// T = Float32 | Float64 (MountainCarEnv.jl:67: the default is Float64); discrete actions (1..3) or, for ContinuousMountainCarEnv,
// a force of type T in -1.0..1.0.
template <class T_> struct MountainCarT {
    using T = T_;
    T action_f = 0;
    static constexpr int NS = 2, NOBS = 2, NACT = 3;
    T min_pos, max_pos, max_speed, goal_pos, goal_velocity, power, gravity;
    int64_t max_steps;
    std::vector<T> state;
    int64_t action = 0;
    bool done = false;
    int64_t t = 0;
    jl::Xoshiro rng;

    MountainCarT(const MountainCarParams& p, jl::Xoshiro r, bool do_reset = true) : state(2, (T)0), rng(r) {
        min_pos = (T)p.min_pos; max_pos = (T)p.max_pos; max_speed = (T)p.max_speed;
        goal_pos = (T)p.goal_pos; goal_velocity = (T)p.goal_velocity; power = (T)p.power;
        gravity = (T)p.gravity; max_steps = p.max_steps;
        if (do_reset) reset();  // MountainCarEnv.jl:79
    }
    void reset() {  // MountainCarEnv.jl:99-105 (0.2 and 0.6 are Float64 literals)
        T u = jl::rand_scalar<T>(rng);
        state[0] = (T)(0.2 * (double)u - 0.6);
        state[1] = 0;
        done = false; t = 0;
    }
    bool act(int64_t a) {  // MountainCarEnv.jl:113-117
        if (a < 1 || a > 3) return false;
        action = a;
        step((int)a - 2);
        return true;
    }
    bool act_continuous(T a) {  // MountainCarEnv.jl:107-111 with a::T: all-T arithmetic
        if (!(a >= (T)-1 && a <= (T)1)) return false;
        action_f = a;
        step_force(a);
        return true;
    }
    void step(int force) { step_force((T)force); }
    void step_force(T force) {  // MountainCarEnv.jl:119-135
        t += 1;
        T x = state[0], v = state[1];
        v = v + (force * power + jl::jcos((T)3 * x) * (-gravity));
        v = jl::jclamp(v, -max_speed, max_speed);
        x = x + v;
        x = jl::jclamp(x, min_pos, max_pos);
        if (x == min_pos && v < 0) v = 0;
        done = (x >= goal_pos && v >= goal_velocity) || t >= max_steps;
        state[0] = x; state[1] = v;
    }
    T reward() const { return done ? (T)0 : (T)-1; }  // MountainCarEnv.jl:95
    void obs(T* out) const { out[0] = state[0]; out[1] = state[1]; }
};
using MountainCar = MountainCarT<float>;

// ------------------------------------------------------------------ Acrobot ----------
// AcrobotEnv (RLEnvs/src/environments/3rd_party/AcrobotEnv.jl:19-225), T = Float64 (the constructor's default, :20).
// reset! :100-107, act! :110-140, dsdt :142-196 (the `book` and `nips` variants), wrap / bound :201-225.
// DEVIATION (documented in DESIGN.md §3): the reference integrates [0, dt] with OrdinaryDiffEq.solve(ode, RK4()) — an adaptive
// step-size controller that lives in an external package absent from the tree; this restatement takes ONE classical RK4 step
// over [0, dt], which is what the file's own source of the equations does ("governing equations as per python gym": gym's
// rk4(derivs, y0, [0, dt])).  Trajectories agree to the solver tolerance, not bit for bit: PARITY UNPINNED.
struct AcrobotParams {  // AcrobotEnvParams{T}
    double link_length_a, link_length_b, link_mass_a, link_mass_b, link_com_pos_a, link_com_pos_b, link_moi, max_torque_noise,
        max_vel_a, max_vel_b, g, dt;
    int64_t max_steps;
    int32_t book;   // book_or_nips == "book" (default) | "nips"
};
inline AcrobotParams acrobot_default_params() {
    return AcrobotParams{1.0, 1.0, 1.0, 1.0, 0.5, 0.5, 1.0, 0.0, 4 * jl::PI_D, 9 * jl::PI_D, 9.8, 0.2, 200, 1};
}
struct Acrobot {
    using T = double;
    static constexpr int NS = 4, NOBS = 6, NACT = 3;
    AcrobotParams p;
    std::vector<T> state;
    int64_t action = 2;
    bool done = false;
    int64_t t = 0;
    T rew = -1;
    jl::Xoshiro rng;
    Acrobot(const AcrobotParams& q, jl::Xoshiro r, bool do_reset = true) : p(q), state(4, 0.0), rng(r) {
        if (do_reset) reset();   // AcrobotEnv.jl:71
    }
    void reset() {  // AcrobotEnv.jl:100-107
        T u[4];
        jl::rand_array4(rng, u);   // rand(env.rng, T, 4)
        for (int k = 0; k < 4; ++k) state[k] = (T)0.1 * u[k] - (T)0.05;
        t = 0; action = 2; done = false; rew = -1;
    }
    // dsdt (AcrobotEnv.jl:142-196); a = the torque carried in the augmented state
    void dsdt(const T* s, T a, T* du) const {
        const T m1 = p.link_mass_a, m2 = p.link_mass_b, l1 = p.link_length_a, lc1 = p.link_com_pos_a, lc2 = p.link_com_pos_b;
        const T I1 = p.link_moi, I2 = p.link_moi, g = p.g;
        const T theta1 = s[0], theta2 = s[1], dtheta1 = s[2], dtheta2 = s[3];
        T ddtheta1 = 0.0, ddtheta2 = 0.0;
        const T c2 = jl::jcos(theta2), s2 = jl::jsin(theta2);
        const T d1 = ((m1 * (lc1 * lc1) + m2 * (((l1 * l1) + (lc2 * lc2)) + ((2 * l1) * lc2) * c2)) + I1) + I2;
        const T d2 = m2 * ((lc2 * lc2) + (l1 * lc2) * c2) + I2;
        const T phi2 = ((m2 * lc2) * g) * jl::jcos((theta1 + theta2) - jl::PI_D / 2.0);
        const T phi1 = (((((((-m2) * l1) * lc2) * (dtheta2 * dtheta2)) * s2) - ((((((2 * m2) * l1) * lc2) * dtheta2) * dtheta1) * s2)) +
                        ((m1 * lc1 + m2 * l1) * g) * jl::jcos(theta1 - jl::PI_D / 2)) + phi2;
        if (!p.book) {
            ddtheta2 = ((a + (d2 / d1) * phi1) - phi2) / (((m2 * (lc2 * lc2)) + I2) - (d2 * d2) / d1);
        } else {
            ddtheta2 = (((a + (d2 / d1) * phi1) - ((((m2 * l1) * lc2) * (dtheta1 * dtheta1)) * s2)) - phi2) / (((m2 * (lc2 * lc2)) + I2) - (d2 * d2) / d1);
            ddtheta1 = (-(d2 * ddtheta2 + phi1)) / d1;
        }
        du[0] = dtheta1; du[1] = dtheta2; du[2] = ddtheta1; du[3] = ddtheta2;
    }
    static T wrap(T x, T m, T M) {   // AcrobotEnv.jl:201-217
        const T diff = M - m;
        while (x > M) x = x - diff;
        while (x < m) x = x + diff;
        return x;
    }
    static T bound(T x, T m, T M) { return std::fmin(std::fmax(x, m), M); }   // AcrobotEnv.jl:219-225 (min(max(x, m), M))
    bool act(int64_t a) {   // AcrobotEnv.jl:110-140
        if (a < 1 || a > 3) return false;   // action_space = Base.OneTo(3)
        action = a;
        t += 1;
        const T torque = (T)(a - 2);         // avail_torque = [-1, 0, 1]  (max_torque_noise = 0: no draw)
        // one classical RK4 step over [0, dt] (see DEVIATION above)
        const T h = p.dt, h2 = p.dt / 2.0;
        T y0[4] = {state[0], state[1], state[2], state[3]}, k1[4], k2[4], k3[4], k4[4], y[4];
        dsdt(y0, torque, k1);
        for (int i = 0; i < 4; ++i) y[i] = y0[i] + h2 * k1[i];
        dsdt(y, torque, k2);
        for (int i = 0; i < 4; ++i) y[i] = y0[i] + h2 * k2[i];
        dsdt(y, torque, k3);
        for (int i = 0; i < 4; ++i) y[i] = y0[i] + h * k3[i];
        dsdt(y, torque, k4);
        T ns[4];
        for (int i = 0; i < 4; ++i) ns[i] = y0[i] + (h / 6.0) * (((k1[i] + 2 * k2[i]) + 2 * k3[i]) + k4[i]);
        ns[0] = wrap(ns[0], -jl::PI_D, jl::PI_D);
        ns[1] = wrap(ns[1], -jl::PI_D, jl::PI_D);
        ns[2] = bound(ns[2], -p.max_vel_a, p.max_vel_a);
        ns[3] = bound(ns[3], -p.max_vel_b, p.max_vel_b);
        for (int i = 0; i < 4; ++i) state[i] = ns[i];
        const bool succeeded = (-jl::jcos(ns[0]) - jl::jcos(ns[1] + ns[0])) > 1.0;
        done = succeeded || t > p.max_steps;
        rew = succeeded ? 0.0 : -1.0;
        return true;
    }
    T reward() const { return rew; }
    void obs(T* out) const {   // acrobot_observation (AcrobotEnv.jl:76)
        out[0] = jl::jcos(state[0]); out[1] = jl::jsin(state[0]); out[2] = jl::jcos(state[1]); out[3] = jl::jsin(state[1]);
        out[4] = state[2]; out[5] = state[3];
    }
};

}  // namespace oracle
