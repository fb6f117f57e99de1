// common.cuh — shared host-side plumbing of libb200rl.so (ctx, error convention, handles).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200rl.h"

#define B200RL_ABI_VERSION 1

void b200rl_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            b200rl_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return _e == cudaErrorMemoryAllocation ? B200RL_ERR_OOM : B200RL_ERR_CUDA;      \
        }                                                                                   \
    } while (0)

#define REQUIRE(cond, code, msg)                 \
    do {                                         \
        if (!(cond)) {                           \
            b200rl_set_error("%s: %s", __func__, msg); \
            return code;                         \
        }                                        \
    } while (0)

#define TRY(expr)                \
    do {                         \
        int _s = (expr);         \
        if (_s != B200RL_OK) return _s; \
    } while (0)

struct b200rl_comm_state;  // comm.cu

struct b200rl_ctx {
    int device = 0;
    int sm_count = 0;
    size_t l2_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    void* scratch = nullptr;  // general device scratch (grown on demand)
    size_t scratch_bytes = 0;
    std::vector<void*> retired;  // outgrown scratch buffers, freed with the ctx
    b200rl_comm_state* comm = nullptr;
    uint64_t launches = 0;    // kernels launched through this ctx (bench "gpu_launches")
};

int ctx_scratch(b200rl_ctx* ctx, size_t bytes, void** out);
static inline int ctx_bind(b200rl_ctx* ctx) {
    if (!ctx) { b200rl_set_error("null ctx"); return B200RL_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    return B200RL_OK;
}
#define LAUNCH_CHECK(ctx)                 \
    do {                                  \
        (ctx)->launches++;                \
        CUDA_TRY(cudaGetLastError());     \
    } while (0)

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// ---- NVLink peer exchange region (comm.cu): one per rank, mapped into every rank of the node ------------------
// Layout (two slots each, slot = sequence number & 1, so a slot is only rewritten after every peer has published the
// NEXT sequence number, i.e. has finished reading it):
//   x[2][kP2PXCap]        fp32   gradient (+ 4 loss sums) published by the fused reduce / exchange / clip / Adam kernel
//   gflag[2][kP2PMaxCta]  u32    per-CTA "chunk published" sequence numbers
//   y[2][kP2PYCap]        f64    small all-reduces (advantage-normalisation sums, ...)
//   yflag[2][16]          u32
constexpr size_t kP2PXCap = 262144;
constexpr int kP2PMaxCta = 1024;
constexpr int kP2PYCap = 1024;
constexpr int kP2PMaxRanks = 8;
constexpr size_t kP2POffG = 2 * kP2PXCap * 4;
constexpr size_t kP2POffY = kP2POffG + 2 * (size_t)kP2PMaxCta * 4;
constexpr size_t kP2POffYF = kP2POffY + 2 * (size_t)kP2PYCap * 8;
constexpr size_t kP2PRegionBytes = kP2POffYF + 2 * 16 * 4;
struct P2PTable {   // nranks == 0: not attached
    int nranks, rank;
    unsigned char* base[kP2PMaxRanks];
};
bool b200rl_comm_p2p_table(b200rl_ctx* ctx, P2PTable* out);   // false when no peer exchange is attached
uint32_t b200rl_comm_p2p_next_gseq(b200rl_ctx* ctx);

#ifdef __CUDACC__
__device__ __forceinline__ float* p2p_x(const P2PTable& t, int r, unsigned slot) { return reinterpret_cast<float*>(t.base[r]) + (size_t)slot * kP2PXCap; }
__device__ __forceinline__ unsigned* p2p_gflag(const P2PTable& t, int r, unsigned slot) {
    return reinterpret_cast<unsigned*>(t.base[r] + kP2POffG) + (size_t)slot * kP2PMaxCta;
}
__device__ __forceinline__ double* p2p_y(const P2PTable& t, int r, unsigned slot) { return reinterpret_cast<double*>(t.base[r] + kP2POffY) + (size_t)slot * kP2PYCap; }
__device__ __forceinline__ unsigned* p2p_yflag(const P2PTable& t, int r, unsigned slot) { return reinterpret_cast<unsigned*>(t.base[r] + kP2POffYF) + (size_t)slot * 16; }
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer_f32(const float* p) {   // peer memory over NVLink: never from a stale cache line
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// spin until the peer's flag carries `seq`; a peer that never arrives (crashed rank) traps after ~10 s instead of hanging the GPU
__device__ __forceinline__ void p2p_wait_flag(const unsigned* flag, unsigned seq) {
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_sys(flag) != seq) {
        if ((++spins & 1023u) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ull) __trap();
        }
    }
}
#endif
