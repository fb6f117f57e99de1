// jl_device.cuh — device-side Julia numerics for the env kernels.
//
// The reference's envs call Julia Base sin/cos/mod and stdlib Random (Xoshiro256++);
// to reproduce its trajectories bit-for-bit the kernels evaluate the same polynomial
// kernels (Julia base/special/trig.jl = FreeBSD msun k_sinf/k_cosf/k_sin/k_cos) in
// Float64 with the same rounding points.  `muladd` sites are explicit __fma_rn; every
// translation unit that includes this header is compiled with -fmad=false so nothing else
// is contracted (Julia/LLVM never contracts a*b+c on its own).
// Reference call sites: CartPoleEnv.jl:122-123, PendulumEnv.jl:70-71,108, MountainCarEnv.jl:122.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace jld {

__device__ __forceinline__ double muladd(double a, double b, double c) { return __fma_rn(a, b, c); }

#define JLD_PI 3.141592653589793

__device__ __forceinline__ float sin_kernel32(double y) {
    double z = y * y;
    double w = z * z;
    double r = muladd(z, 2.718311493989822e-6, -0.00019839334836096632);
    double s = z * y;
    return (float)((y + s * muladd(z, 0.008333329385889463, -0.16666666641626524)) + (s * w) * r);
}
__device__ __forceinline__ float cos_kernel32(double y) {
    double z = y * y;
    double w = z * z;
    double r = muladd(z, 2.439044879627741e-5, -0.001388676377460993);
    return (float)(((1.0 + z * -0.499999997251031) + w * 0.04166662332373906) + (w * z) * r);
}
// Cody–Waite reduction of a Float32 argument (Julia rem_pio2_kernel(::Float32)), valid for
// |x| < Float32(pi)/2 * 2^28; larger arguments are unreachable for the envs.
__device__ __forceinline__ int rem_pio2_32(float x, double* y) {
    double xd = (double)x;
    double ax = fabs(xd);
    if (ax <= JLD_PI * 5 / 4) {
        if (ax <= JLD_PI * 3 / 4) {
            if (x > 0) { *y = xd - JLD_PI / 2; return 1; }
            *y = xd + JLD_PI / 2; return -1;
        }
        if (x > 0) { *y = xd - JLD_PI; return 2; }
        *y = xd + JLD_PI; return -2;
    } else if (ax <= JLD_PI * 9 / 4) {
        if (ax <= JLD_PI * 7 / 4) {
            if (x > 0) { *y = xd - JLD_PI * 3 / 2; return 3; }
            *y = xd + JLD_PI * 3 / 2; return -3;
        }
        if (x > 0) { *y = xd - JLD_PI * 4 / 2; return 4; }
        *y = xd + JLD_PI * 4 / 2; return -4;
    }
    double fn = rint(xd * 6.36619772367581382433e-01);
    double r = xd - fn * 1.57079631090164184570e+00;
    double w = fn * 1.58932547735281966916e-08;
    *y = r - w;
    return (int)(long long)fn;
}
__device__ __forceinline__ float jsin(float x) {
    float ax = fabsf(x);
    if (ax < 0.78539819f) {
        if (ax < 0x1.6a09e6p-12f) return x;
        return sin_kernel32((double)x);
    }
    if (!(ax < __int_as_float(0x7f800000))) return __int_as_float(0x7fc00000);
    double y;
    int n = rem_pio2_32(x, &y) & 3;
    if (n == 0) return sin_kernel32(y);
    if (n == 1) return cos_kernel32(y);
    if (n == 2) return -sin_kernel32(y);
    return -cos_kernel32(y);
}
__device__ __forceinline__ float jcos(float x) {
    float ax = fabsf(x);
    if (ax < 0.78539819f) {
        if (ax < 0x1p-12f) return 1.0f;
        return cos_kernel32((double)x);
    }
    if (!(ax < __int_as_float(0x7f800000))) return __int_as_float(0x7fc00000);
    double y;
    int n = rem_pio2_32(x, &y) & 3;
    if (n == 0) return cos_kernel32(y);
    if (n == 1) return -sin_kernel32(y);
    if (n == 2) return -cos_kernel32(y);
    return sin_kernel32(y);
}

// ---- Float64 (CartPoleEnv{Float64}, BASELINE config 1) --------------------------------
#define DS1 -1.66666666666666324348e-01
#define DS2 8.33333333332248946124e-03
#define DS3 -1.98412698298579493134e-04
#define DS4 2.75573137070700676789e-06
#define DS5 -2.50507602534068634195e-08
#define DS6 1.58969099521155010221e-10
#define DC1 4.16666666666666019037e-02
#define DC2 -1.38888888888741095749e-03
#define DC3 2.48015872894767294178e-05
#define DC4 -2.75573143513906633035e-07
#define DC5 2.08757232129817482790e-09
#define DC6 -1.13596475577881948265e-11
__device__ __forceinline__ double sin_kernel64(double y) {
    double y2 = y * y, y4 = y2 * y2;
    double r = muladd(y2, muladd(y2, DS4, DS3), DS2) + y2 * y4 * muladd(y2, DS6, DS5);
    double y3 = y2 * y;
    return y + y3 * (DS1 + y2 * r);
}
__device__ __forceinline__ double sin_kernel64(double hi, double lo) {
    double y2 = hi * hi, y4 = y2 * y2;
    double r = muladd(y2, muladd(y2, DS4, DS3), DS2) + y2 * y4 * muladd(y2, DS6, DS5);
    double y3 = y2 * hi;
    return hi - ((y2 * (0.5 * lo - y3 * r) - lo) - y3 * DS1);
}
__device__ __forceinline__ double cos_kernel64(double hi, double lo) {
    double y2 = hi * hi, y4 = y2 * y2;
    double r = y2 * muladd(y2, muladd(y2, DC3, DC2), DC1) + y4 * y4 * muladd(y2, muladd(y2, DC6, DC5), DC4);
    double half = 0.5 * y2;
    double w = 1.0 - half;
    return w + (((1.0 - w) - half) + (y2 * r - hi * lo));
}
// cody_waite_2c_pio2(x, fn, n) (rem_pio2.jl): two-constant reduction used for |x| <= 9pi/4 away from multiples of pi/2
__device__ __forceinline__ int cody_waite_2c(double x, double fn, int n, double* y1o, double* y2o) {
    double z = muladd(-fn, 1.57079632673412561417e+00, x);
    double w = fn * 6.07710050650619224932e-11;
    double y1 = z - w;
    *y1o = y1;
    *y2o = (z - y1) - w;
    return n;
}
// cody_waite_ext_pio2 (medium range, |x| < 2^20*pi/2): up to three rounds
__device__ __forceinline__ int cody_waite_ext(double x, unsigned xhp, double* y1o, double* y2o) {
    double fn = rint(x * 6.36619772367581382433e-01);
    double r = muladd(-fn, 1.57079632673412561417e+00, x);
    double w = fn * 6.07710050650619224932e-11;
    int j = (int)(xhp >> 20);
    double y1 = r - w;
    int i = j - (int)(((unsigned)__double2hiint(y1) >> 20) & 0x7ff);
    if (i > 16) {
        double t = r;
        w = fn * 6.07710050630396597660e-11;
        r = t - w;
        w = muladd(fn, 2.02226624879595063154e-21, -((t - r) - w));
        y1 = r - w;
        i = j - (int)(((unsigned)__double2hiint(y1) >> 20) & 0x7ff);
        if (i > 49) {
            t = r;
            w = fn * 2.02226624871116645580e-21;
            r = t - w;
            w = muladd(fn, 8.47842766036889956997e-32, -((t - r) - w));
            y1 = r - w;
        }
    }
    *y1o = y1;
    *y2o = (r - y1) - w;
    return (int)(long long)fn;
}
// rem_pio2_kernel(x::Float64) (base/special/rem_pio2.jl, a port of msun e_rem_pio2.c): the decision tree on the high word —
// |x| <= 9pi/4 uses the two-constant scheme with fn = +-1..4 unless x is close to a multiple of pi/2, everything up to
// 2^20 pi/2 the extended scheme; Payne-Hanek beyond is not restated (unreachable: Pendulum |theta| < 100, MountainCar |3x| < 4).
__device__ __forceinline__ int rem_pio2_64(double x, double* y1o, double* y2o) {
    const unsigned xhp = (unsigned)__double2hiint(x) & 0x7fffffffu;
    const bool pos = x > 0.0;
    if (xhp <= 0x400f6a7au) {                       // |x| ~<= 5pi/4
        if ((xhp & 0xfffffu) == 0x921fbu) return cody_waite_ext(x, xhp, y1o, y2o);     // |x| ~= pi/2 or 2pi/2
        if (xhp <= 0x4002d97cu) return pos ? cody_waite_2c(x, 1.0, 1, y1o, y2o) : cody_waite_2c(x, -1.0, -1, y1o, y2o);   // |x| ~<= 3pi/4
        return pos ? cody_waite_2c(x, 2.0, 2, y1o, y2o) : cody_waite_2c(x, -2.0, -2, y1o, y2o);
    }
    if (xhp <= 0x401c463bu) {                       // |x| ~<= 9pi/4
        if (xhp <= 0x4015fdbcu) {                   // |x| ~<= 7pi/4
            if (xhp == 0x4012d97cu) return cody_waite_ext(x, xhp, y1o, y2o);           // |x| ~= 3pi/2
            return pos ? cody_waite_2c(x, 3.0, 3, y1o, y2o) : cody_waite_2c(x, -3.0, -3, y1o, y2o);
        }
        if (xhp == 0x401921fbu) return cody_waite_ext(x, xhp, y1o, y2o);               // |x| ~= 4pi/2
        return pos ? cody_waite_2c(x, 4.0, 4, y1o, y2o) : cody_waite_2c(x, -4.0, -4, y1o, y2o);
    }
    return cody_waite_ext(x, xhp, y1o, y2o);
}
__device__ __forceinline__ double jsin(double x) {
    double ax = fabs(x);
    if (ax < JLD_PI / 4) {
        if (ax < 0x1p-26) return x;
        return sin_kernel64(x);
    }
    if (!(ax < __longlong_as_double(0x7ff0000000000000ll))) return __longlong_as_double(0x7ff8000000000000ll);
    double hi, lo;
    int n = rem_pio2_64(x, &hi, &lo) & 3;
    if (n == 0) return sin_kernel64(hi, lo);
    if (n == 1) return cos_kernel64(hi, lo);
    if (n == 2) return -sin_kernel64(hi, lo);
    return -cos_kernel64(hi, lo);
}
__device__ __forceinline__ double jcos(double x) {
    double ax = fabs(x);
    if (ax < JLD_PI / 4) {
        if (ax < 0x1.6a09e667f3bcdp-27) return 1.0;
        return cos_kernel64(x, 0.0);
    }
    if (!(ax < __longlong_as_double(0x7ff0000000000000ll))) return __longlong_as_double(0x7ff8000000000000ll);
    double hi, lo;
    int n = rem_pio2_64(x, &hi, &lo) & 3;
    if (n == 0) return cos_kernel64(hi, lo);
    if (n == 1) return -sin_kernel64(hi, lo);
    if (n == 2) return -cos_kernel64(hi, lo);
    return sin_kernel64(hi, lo);
}

// Base.mod(::Float64, ::Float64)
__device__ __forceinline__ double jmod(double x, double y) {
    double r = fmod(x, y);
    if (r == 0) return copysign(r, y);
    if ((r > 0) != (y > 0)) return r + y;
    return r;
}
template <class T> __device__ __forceinline__ T jclamp(T x, T lo, T hi) { return x > hi ? hi : (x < lo ? lo : x); }
// x * b::Bool — Julia's strong zero
__device__ __forceinline__ float mul_bool(float x, bool b) { return b ? x : copysignf(0.0f, x); }
__device__ __forceinline__ double mul_bool(double x, bool b) { return b ? x : copysign(0.0, x); }

// ---- Xoshiro256++ and the Julia 1.10 samplers (stdlib Random) ----------
struct Xo { unsigned long long s0, s1, s2, s3; };
__device__ __forceinline__ unsigned long long rotl(unsigned long long x, int k) { return (x << k) | (x >> (64 - k)); }
__device__ __forceinline__ unsigned long long next(Xo& g) {
    unsigned long long s0 = g.s0, s1 = g.s1, s2 = g.s2, s3 = g.s3;
    unsigned long long res = rotl(s0 + s3, 23) + s0;
    unsigned long long t = s1 << 17;
    s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t;
    s3 = rotl(s3, 45);
    g.s0 = s0; g.s1 = s1; g.s2 = s2; g.s3 = s3;
    return res;
}
__device__ __forceinline__ double rand_f64(Xo& g) { return (double)(next(g) >> 11) * 0x1p-53; }
__device__ __forceinline__ float rand_f32(Xo& g) { return (float)((unsigned)(next(g) >> 32) >> 8) * 0x1p-24f; }
template <class T> __device__ __forceinline__ T rand_real(Xo& g);     // rand(rng, T), scalar API
template <> __device__ __forceinline__ float rand_real<float>(Xo& g) { return rand_f32(g); }
template <> __device__ __forceinline__ double rand_real<double>(Xo& g) { return rand_f64(g); }
// rand(rng, T, 4): array API below the 64-byte SIMD threshold (one u64 per 8 output bytes)
__device__ __forceinline__ void rand4(Xo& g, float* o) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        unsigned long long u = next(g);
        o[2 * k] = (float)((unsigned)u >> 8) * 0x1p-24f;
        o[2 * k + 1] = (float)((unsigned)(u >> 32) >> 8) * 0x1p-24f;
    }
}
__device__ __forceinline__ void rand4(Xo& g, double* o) {
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = rand_f64(g);
}
// rand(rng, Base.OneTo(n)) — Lemire nearly-divisionless on UInt64 (SamplerRangeNDL)
__device__ __forceinline__ long long rand_oneto(Xo& g, unsigned long long n) {
    unsigned long long x = next(g);
    unsigned long long hi = __umul64hi(x, n), lo = x * n;
    if (lo < n) {
        unsigned long long t = (0ull - n) % n;
        while (lo < t) {
            x = next(g);
            hi = __umul64hi(x, n);
            lo = x * n;
        }
    }
    return (long long)hi + 1;
}

}  // namespace jld
