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
