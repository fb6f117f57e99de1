// ORACLE — TEST INFRASTRUCTURE ONLY (see jl_math.hpp header).
// Flat C entry points over the oracle headers so tests/ and bench.py's cpu_baseline /
// --impl reference legs can drive it through ctypes.  Nothing in the product path links
// or loads this library.
#include <chrono>
#include <cstdint>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "envs.hpp"
#include "returns.hpp"
#include "vecenv.hpp"
#include "traj.hpp"
#include "nn.hpp"

using namespace oracle;

extern "C" {

int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// ---- scalar math / rng probes (known-answer tests) -----------------------------------
float orc_sin32(float x) { return jl::sin32(x); }
float orc_cos32(float x) { return jl::cos32(x); }
double orc_sin64(double x) { return jl::sin64(x); }
double orc_cos64(double x) { return jl::cos64(x); }
double orc_mod64(double x, double y) { return jl::jmod(x, y); }
uint64_t orc_rng_next(uint64_t* s) {
    jl::Xoshiro g{s[0], s[1], s[2], s[3]};
    uint64_t r = jl::next_u64(g);
    s[0] = g.s0; s[1] = g.s1; s[2] = g.s2; s[3] = g.s3;
    return r;
}
int64_t orc_rng_oneto(uint64_t* s, uint64_t n) {
    jl::Xoshiro g{s[0], s[1], s[2], s[3]};
    int64_t r = jl::rand_oneto(g, n);
    s[0] = g.s0; s[1] = g.s1; s[2] = g.s2; s[3] = g.s3;
    return r;
}
// shuffle!(rng, a) / randperm(rng, n) of Julia's stdlib on a raw Xoshiro state (advanced in place)
void orc_jl_shuffle_i64(uint64_t* s, int64_t* a, int64_t n) {
    jl::Xoshiro g{s[0], s[1], s[2], s[3]};
    jl::shuffle(g, a, n);
    s[0] = g.s0; s[1] = g.s1; s[2] = g.s2; s[3] = g.s3;
}
void orc_jl_shuffle_i32(uint64_t* s, int32_t* a, int64_t n) {
    jl::Xoshiro g{s[0], s[1], s[2], s[3]};
    jl::shuffle(g, a, n);
    s[0] = g.s0; s[1] = g.s1; s[2] = g.s2; s[3] = g.s3;
}
void orc_jl_randperm(uint64_t* s, int64_t* a, int64_t n) {
    jl::Xoshiro g{s[0], s[1], s[2], s[3]};
    jl::randperm(g, a, n);
    s[0] = g.s0; s[1] = g.s1; s[2] = g.s2; s[3] = g.s3;
}
uint64_t orc_jl_ltm52(uint64_t* s, uint64_t n, uint64_t mask) {
    jl::Xoshiro g{s[0], s[1], s[2], s[3]};
    uint64_t r = jl::rand_ltm52(g, n, mask);
    s[0] = g.s0; s[1] = g.s1; s[2] = g.s2; s[3] = g.s3;
    return r;
}
void orc_seed_splitmix(uint64_t seed, uint64_t* out) {
    jl::Xoshiro g = jl::seed_splitmix(seed);
    out[0] = g.s0; out[1] = g.s1; out[2] = g.s2; out[3] = g.s3;
}

// ---- default params (what the reference constructors produce) ------------------------
void orc_cartpole_default_params(int dtype, double* out /*11*/) {
    CartPoleParams p = dtype == 1 ? cartpole_default_params<double>() : cartpole_default_params<float>();
    double v[11] = {p.gravity, p.masscart, p.masspole, p.totalmass, p.halflength, p.polemasslength,
                    p.forcemag, p.dt, p.thetathreshold, p.xthreshold, (double)p.max_steps};
    std::memcpy(out, v, sizeof v);
}

// ---- vector env ----------------------------------------------------------------------
// kind: 0 CartPole, 1 Pendulum, 2 MountainCar, 3 continuous CartPole, 4 ContinuousMountainCar; dtype: 0 f32, 1 f64 (not for the continuous CartPole)
// params: CartPole 11 doubles (above order); Pendulum 9 (max_speed,max_torque,g,m,l,dt,
// max_steps,n_actions,continuous); MountainCar 8 (min_pos,max_pos,max_speed,goal_pos,
// goal_velocity,power,gravity,max_steps).
void* orc_vecenv_create(int kind, int dtype, int64_t N, const double* q, const uint64_t* rng) {
    if (kind == 0) {
        CartPoleParams p{q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9], (int64_t)q[10]};
        if (dtype == 1) return (VecEnvBase*)new VecCartPoleF64(N, p, rng);
        return (VecEnvBase*)new VecCartPoleF32(N, p, rng);
    } else if (kind == 1) {
        PendulumParams p{q[0], q[1], q[2], q[3], q[4], q[5], (int64_t)q[6], (int64_t)q[7], (int32_t)q[8]};
        if (dtype == 1) return p.continuous ? (VecEnvBase*)new VecPendulum64C(N, p, rng) : (VecEnvBase*)new VecPendulum64D(N, p, rng);
        if (p.continuous) return (VecEnvBase*)new VecPendulumC(N, p, rng);
        return (VecEnvBase*)new VecPendulumD(N, p, rng);
    } else if (kind == 2) {
        MountainCarParams p{q[0], q[1], q[2], q[3], q[4], q[5], q[6], (int64_t)q[7]};
        if (dtype == 1) return (VecEnvBase*)new VecMountainCar64(N, p, rng);
        return (VecEnvBase*)new VecMountainCar(N, p, rng);
    } else if (kind == 3) {   // CartPoleEnv(continuous = true), Float32
        CartPoleParams p{q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9], (int64_t)q[10]};
        return (VecEnvBase*)new VecCartPoleC(N, p, rng);
    } else if (kind == 4) {   // ContinuousMountainCarEnv, Float32
        MountainCarParams p{q[0], q[1], q[2], q[3], q[4], q[5], q[6], (int64_t)q[7]};
        if (dtype == 1) return (VecEnvBase*)new VecMountainCar64C(N, p, rng);
        return (VecEnvBase*)new VecMountainCarC(N, p, rng);
    } else if (kind == 5) {   // AcrobotEnv{Float64}: 14 doubles in AcrobotParams order
        if (dtype != 1) return nullptr;
        AcrobotParams p{q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9], q[10], q[11], (int64_t)q[12], (int32_t)q[13]};
        return (VecEnvBase*)new VecAcrobot(N, p, rng);
    }
    return nullptr;
}
void orc_vecenv_set_max_timeout(void* h, int64_t max_t) { ((VecEnvBase*)h)->max_timeout = max_t; }
void orc_vecenv_destroy(void* h) { delete (VecEnvBase*)h; }
void orc_vecenv_reset(void* h, int force) { ((VecEnvBase*)h)->reset(force); }
int orc_vecenv_step(void* h, const void* actions, int auto_reset) { return ((VecEnvBase*)h)->step(actions, auto_reset); }
int orc_vecenv_step_random(void* h, int auto_reset, int32_t* actions_out) { return ((VecEnvBase*)h)->step_random(auto_reset, actions_out); }
void orc_vecenv_get(void* h, int field, void* dst) { ((VecEnvBase*)h)->get(field, dst); }
void orc_vecenv_set(void* h, int field, const void* src) { ((VecEnvBase*)h)->set(field, src); }

// CPU baseline: `steps` random-policy steps with auto-reset; returns seconds.
double orc_vecenv_bench_random(void* h, int steps) {
    VecEnvBase* e = (VecEnvBase*)h;
    auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < steps; ++s) e->step_random(1, nullptr);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- returns -------------------------------------------------------------------------
void orc_gae_f32(float* adv, const float* r, const float* v, const uint8_t* term, float gamma, float lambda, int64_t R, int64_t C, int dims) { gae_mat<float>(adv, r, v, term, gamma, lambda, R, C, dims); }
void orc_gae_f64(double* adv, const double* r, const double* v, const uint8_t* term, double gamma, double lambda, int64_t R, int64_t C, int dims) { gae_mat<double>(adv, r, v, term, gamma, lambda, R, C, dims); }
void orc_discount_f32(float* out, const float* r, const uint8_t* term, const float* init, float gamma, int64_t R, int64_t C, int dims) { discount_rewards_mat<float>(out, r, term, init, gamma, R, C, dims); }
void orc_discount_f64(double* out, const double* r, const uint8_t* term, const double* init, double gamma, int64_t R, int64_t C, int dims) { discount_rewards_mat<double>(out, r, term, init, gamma, R, C, dims); }
void orc_discount_reduced_f32(float* out, const float* r, const uint8_t* term, const float* init, float gamma, int64_t R, int64_t C, int dims) { discount_rewards_reduced_mat<float>(out, r, term, init, gamma, R, C, dims); }
void orc_discount_reduced_f64(double* out, const double* r, const uint8_t* term, const double* init, double gamma, int64_t R, int64_t C, int dims) { discount_rewards_reduced_mat<double>(out, r, term, init, gamma, R, C, dims); }

}  // extern "C"

#include "capi_traj_nn.inc"
