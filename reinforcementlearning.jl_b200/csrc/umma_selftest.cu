// umma_selftest.cu — NOT part of libb200rl.so: built into build/libb200rl_selftest.so by `build.py --selftest` for the probes under
// profiles/ (umma_probe*.py).  Diagnostic entry point that pins down the tcgen05 operand-layout / TMEM
// mapping conventions on real hardware: the host supplies raw shared-memory images of the A and B
// operands plus descriptor parameters; the kernel issues the MMAs and dumps TMEM lanes 0..127.
#include "common.cuh"
#include "umma.cuh"

namespace {
struct SelfTestArgs {
    const uint8_t* a_img; uint32_t a_bytes;
    const uint8_t* b_img; uint32_t b_bytes;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;      // descriptor byte offsets
    uint32_t a_kadv, b_kadv;                  // start-address advance per K step (bytes)
    uint32_t idesc; int ksteps; int ncols;    // instruction descriptor, # MMA instructions, D columns to dump
    int a_from_tmem;                           // 1: copy the (K-major, 64-wide) A image into TMEM columns 64.. and use the .ts form
    int f16;                                   // 1: kind::f16 instead of kind::tf32
    int a_raw;                                 // 1 (with a_from_tmem): the A image is a raw row-major [128][64] table of 32-bit TMEM words
    int repeat;                                // timing: issue the whole k-loop this many times; cycles -> d_out[last]
    float* d_out;                             // [128][ncols]
};

__global__ void __launch_bounds__(128) umma_selftest_kernel(SelfTestArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((a.a_bytes + 1023) / 1024) * 1024;
    for (uint32_t k = threadIdx.x; k < a.a_bytes / 4; k += blockDim.x) ((uint32_t*)sa)[k] = ((const uint32_t*)a.a_img)[k];
    for (uint32_t k = threadIdx.x; k < a.b_bytes / 4; k += blockDim.x) ((uint32_t*)sb)[k] = ((const uint32_t*)a.b_img)[k];
    umma::fence_proxy_async();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
    if (threadIdx.x == 32) umma::mbar_init(&s_bar, 1);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;
    if (a.a_from_tmem) {  // thread t <-> row m = t: A[m][k] (k < 64) read from the K-major image, stored to TMEM lane m, column 64 + k
        const int m = threadIdx.x;
        for (int c0 = 0; c0 < 64; c0 += 16) {
            float v[16];
            for (int k = 0; k < 16; ++k) {
                int kk = c0 + k;
                v[k] = a.a_raw ? reinterpret_cast<const float*>(sa)[m * 64 + kk]
                               : *reinterpret_cast<const float*>(sa + (m / 8) * a.a_sbo + (kk / 4) * a.a_lbo + (m % 8) * 16 + (kk % 4) * 4);
            }
            umma::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
        }
        umma::tmem_st_wait();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
    }
    long long t0 = clock64();
    if (warp == 0) {
        if (threadIdx.x == 0) {
            for (int rep = 0; rep < a.repeat; ++rep)
            for (int k = 0; k < a.ksteps; ++k) {
                uint64_t db = umma::make_desc(umma::smem_u32(sb) + k * a.b_kadv, a.b_lbo, a.b_sbo);
                const uint32_t acc = (k > 0 || rep > 0) ? 1u : 0u;
                if (a.a_from_tmem) {
                    if (a.f16) umma::mma_f16_ts(tmem, tmem + 256 + k * a.a_kadv, db, a.idesc, acc);
                    else umma::mma_tf32_ts(tmem, tmem + 256 + k * a.a_kadv, db, a.idesc, acc);
                } else {
                    uint64_t da = umma::make_desc(umma::smem_u32(sa) + k * a.a_kadv, a.a_lbo, a.a_sbo);
                    if (a.f16) umma::mma_f16(tmem, da, db, a.idesc, acc);
                    else umma::mma_tf32(tmem, da, db, a.idesc, acc);
                }
            }
            umma::commit(&s_bar);
        }
        __syncwarp();
    }
    umma::mbar_wait(&s_bar, 0);
    long long t1 = clock64();
    umma::fence_after_sync();
    for (int c0 = 0; c0 < a.ncols; c0 += 32) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        int lane = warp * 32 + (threadIdx.x & 31);
        for (int c = 0; c < 32; ++c) a.d_out[lane * a.ncols + c0 + c] = v[c];
    }
    __syncthreads();
    if (threadIdx.x == 0) a.d_out[128 * a.ncols - 1] = (float)(t1 - t0);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" int b200rl_selftest_umma(b200rl_ctx* ctx, const void* a_img_host, uint32_t a_bytes, const void* b_img_host, uint32_t b_bytes,
                                    const uint32_t* desc8 /* a_lbo,a_sbo,b_lbo,b_sbo,a_kadv,b_kadv,idesc,ksteps | (a_from_tmem << 16) | (f16 << 17) | (a_raw << 18) | (repeat << 20) */, int ncols,
                                    float* d_out_host /* [128][ncols] */) {
    TRY(ctx_bind(ctx));
    REQUIRE(a_img_host && b_img_host && desc8 && d_out_host, B200RL_ERR_INVALID, "null argument");
    REQUIRE(ncols == 32 || ncols == 64 || ncols == 128, B200RL_ERR_INVALID, "ncols must be 32, 64 or 128");
    REQUIRE(a_bytes % 4 == 0 && b_bytes % 4 == 0 && a_bytes + b_bytes < 200 * 1024, B200RL_ERR_INVALID, "bad image sizes");
    void* sc;
    size_t a_pad = ((size_t)a_bytes + 1023) / 1024 * 1024, b_pad = ((size_t)b_bytes + 1023) / 1024 * 1024;
    TRY(ctx_scratch(ctx, a_pad + b_pad + (size_t)128 * ncols * 4 + 1024, &sc));
    uint8_t* da = (uint8_t*)sc; uint8_t* db = da + a_pad; float* dd = (float*)(db + b_pad);
    CUDA_TRY(cudaMemcpyAsync(da, a_img_host, a_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(db, b_img_host, b_bytes, cudaMemcpyHostToDevice, ctx->stream));
    SelfTestArgs a{da, a_bytes, db, b_bytes, desc8[0], desc8[1], desc8[2], desc8[3], desc8[4], desc8[5], desc8[6], (int)(desc8[7] & 0xFFFF), ncols, (int)((desc8[7] >> 16) & 1), (int)((desc8[7] >> 17) & 1), (int)((desc8[7] >> 18) & 1), (int)(desc8[7] >> 20) ? (int)(desc8[7] >> 20) : 1, dd};
    size_t smem = a_pad + b_pad;
    CUDA_TRY(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, ctx->stream>>>(a);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(d_out_host, dd, (size_t)128 * ncols * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200RL_OK;
}

// ---- pacing probe: cycles per tcgen05.mma for a given shape / shared-memory layout type / operand source -------------------
// (operand VALUES are irrelevant: the images are zero-filled; only the issue and completion times are read)
namespace {
struct PaceArgs {
    uint32_t idesc;          // instruction descriptor (M, N, kind::f16)
    uint32_t layout;         // shared-memory descriptor layout_type (bits 61..63): 0 none, 1 128B_base32B, 2 128B, 4 64B, 6 32B
    uint32_t lbo, sbo;       // descriptor byte offsets
    uint32_t kadv;           // start-address advance per instruction (bytes), cycled over 4 steps
    int n_mma;               // instructions between the two clock reads
    int ts;                  // 1: A from TMEM
    int alt_d;               // 1: alternate between two accumulators (columns 0 / 128)
    int sleepers;            // 1: the other threads nanosleep instead of spinning on the mbarrier
    int elect;               // 1: the issuing lane is chosen by elect.sync (0: threadIdx.x == 0)
    float* out;              // [0] issue cycles, [1] issue + completion cycles
};
__global__ void __launch_bounds__(128) umma_pacing_kernel(PaceArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    for (uint32_t k = threadIdx.x; k < 128 * 1024 / 4; k += blockDim.x) ((uint32_t*)smem)[k] = 0u;
    umma::fence_proxy_async();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
    if (threadIdx.x == 32) umma::mbar_init(&s_bar, 1);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint64_t hi = ((uint64_t)(a.layout & 7u) << 61);
    const uint64_t dA = umma::make_desc(umma::smem_u32(smem), a.lbo, a.sbo) | hi;
    const uint64_t dB = umma::make_desc(umma::smem_u32(smem + 64 * 1024), a.lbo, a.sbo) | hi;
    long long t0 = clock64(), t1 = t0;
    if (a.elect) {
        if (warp == 0 && umma::elect_one()) {
            for (int i = 0; i < a.n_mma; i += 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t d = tmem + ((a.alt_d && (j & 1)) ? 128u : 0u);
                    if (a.ts) umma::mma_f16_ts(d, tmem + 448 + 8 * j, dB + (uint64_t)((j * a.kadv) >> 4), a.idesc, i + j > 1 ? 1u : 0u);
                    else umma::mma_f16(d, dA + (uint64_t)((j * a.kadv) >> 4), dB + (uint64_t)((j * a.kadv) >> 4), a.idesc, i + j > 1 ? 1u : 0u);
                }
            }
            t1 = clock64();
            umma::commit(&s_bar);
        }
        __syncwarp();
    } else if (threadIdx.x == 0) {
        for (int i = 0; i < a.n_mma; ++i) {
            const uint64_t adv = (uint64_t)(((i & 3) * a.kadv) >> 4);
            const uint32_t d = tmem + ((a.alt_d && (i & 1)) ? 128u : 0u);
            if (a.ts) umma::mma_f16_ts(d, tmem + 448 + 8 * (i & 3), dB + adv, a.idesc, i > 1 ? 1u : 0u);
            else umma::mma_f16(d, dA + adv, dB + adv, a.idesc, i > 1 ? 1u : 0u);
        }
        t1 = clock64();
        umma::commit(&s_bar);
    }
    if (a.sleepers && threadIdx.x != 0) __nanosleep(20000 + 60 * a.n_mma);
    umma::mbar_wait(&s_bar, 0);
    long long t2 = clock64();
    umma::fence_after_sync();
    if (t1 != t0) { a.out[0] = (float)(t1 - t0); a.out[1] = (float)(t2 - t0); }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" int b200rl_selftest_pacing(b200rl_ctx* ctx, const uint32_t* p10 /* idesc, layout, lbo, sbo, kadv, n_mma, ts, alt_d, sleepers, elect */,
                                      float* out2_host) {
    TRY(ctx_bind(ctx));
    REQUIRE(p10 && out2_host, B200RL_ERR_INVALID, "null argument");
    void* sc;
    TRY(ctx_scratch(ctx, 1024, &sc));
    sc = (void*)(((uintptr_t)sc + 15) & ~(uintptr_t)15);
    PaceArgs a{p10[0], p10[1], p10[2], p10[3], p10[4], (int)p10[5], (int)p10[6], (int)p10[7], (int)p10[8], (int)p10[9], (float*)sc};
    const size_t smem = 128 * 1024;
    CUDA_TRY(cudaFuncSetAttribute(umma_pacing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_pacing_kernel<<<1, 128, smem, ctx->stream>>>(a);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(out2_host, sc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200RL_OK;
}
