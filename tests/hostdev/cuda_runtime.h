// tests/hostdev/cuda_runtime.h — TEST INFRASTRUCTURE.  A stand-in for <cuda_runtime.h> that lets g++ compile the *device* headers of
// the product (csrc/jl_device.cuh, env_device.cuh, perm.cuh) for the host, so the `-m "not gpu"` suite can run the exact source the
// kernels execute against the oracle.  Only what those headers use: the qualifiers, four vector types, five intrinsics.
// IEEE semantics are the same on both sides: g++ is invoked with -ffp-contract=off (the .cu files use -fmad=false), explicit
// muladd sites map to std::fma, conversions and rint round to nearest even.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct ulonglong2 { unsigned long long x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline int __double2hiint(double d) { long long i; std::memcpy(&i, &d, 8); return (int)(i >> 32); }
