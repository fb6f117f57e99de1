// nn_tc.cu — tensor-core (tcgen05 / TMEM) variants of the learner kernels for H = 64.
//
// The 64 x 64 layer is a [128 samples x 64] x [64 x 64] GEMM per tile: one elected thread issues
// tcgen05.mma kind::tf32 with operands in shared memory (SWIZZLE_NONE canonical layout, see
// umma.cuh) and the FP32 accumulator in TMEM; the epilogue warps pull it back with tcgen05.ld.
// Parity needs ~FP32 accuracy (1e-5 relative on losses), which plain TF32 (10-bit mantissa,
// measured 7.8e-4 relative on this GEMM) cannot give, so every product is the 3xTF32 split
//     A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi,   x_hi = x with the low 13 mantissa bits cleared
// (the hardware ignores those bits, so the FULL fp32 image doubles as the hi operand) — measured
// 5e-7 relative (profiles/umma_probe.py).  Layer 1 (K <= 4) and the heads (N <= 4) stay on FFMA.
#include "nn.cuh"
#include "umma.cuh"

namespace {

constexpr int NT = 256;
constexpr int TM = 128;              // samples per tile = UMMA M
constexpr int H = 64;
constexpr int G_F = 128;             // byte stride between 4-feature chunks (K direction)
constexpr int G_S = 16 * G_F + 16;   // byte stride between 8-sample groups (+16: bank spread)
constexpr int IMG_BYTES = 16 * G_S;  // one [128 x 64] activation image
constexpr int GW_S = 16 * G_F;       // weight image: stride between 8-row groups
constexpr int WIMG_BYTES = 8 * GW_S; // one [64 x 64] weight image
constexpr float kLog2Pi = 1.8378770664093453f;

struct SmemFwd {
    alignas(128) uint8_t A_full[IMG_BYTES];   // H1 (fp32; the tensor core reads its tf32 prefix = hi part)
    alignas(128) uint8_t A_lo[IMG_BYTES];     // H1 - hi(H1)
    alignas(128) uint8_t B_full[WIMG_BYTES];  // W2 as (n = out, k = in), K-major
    alignas(128) uint8_t B_lo[WIMG_BYTES];
    float W1[kInMax * H];                      // [i][o]
    float b1[H], b2[H];
    float W3[H * kOutMax];                     // [j][o]
    float b3[kOutMax];
    float X[kInMax * TM];                      // [i][s]
    float Zp[2 * kOutMax * TM];                // head partials [half][o][s]
    alignas(8) uint64_t bar;
    uint32_t tmem;
};

__device__ __forceinline__ float act_f(int act, float z) { return act == B200RL_ACT_RELU ? fmaxf(z, 0.f) : tanhf(z); }
__device__ __forceinline__ float hi_part(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float softplus_f(float x) { return x > 0.f ? x + log1pf(expf(-x)) : log1pf(expf(x)); }
__device__ __forceinline__ float normlogpdf1(float mu, float sigma, float x) {
    float s = sigma + 1e-8f, v = s * s, dd = x - mu;
    return -0.5f * ((logf(v) + (dd * dd) / v) + kLog2Pi);
}
__device__ __forceinline__ int64_t head_base(const MlpDesc& d) { return (int64_t)d.H * d.in + d.H + (int64_t)d.H * d.H + d.H; }
__device__ __forceinline__ int64_t head_w(const MlpDesc& d, int o, int j) {
    return head_base(d) + (d.heads2 ? (int64_t)o * (d.H + 1) + j : (int64_t)o + (int64_t)d.nout * j);
}
__device__ __forceinline__ int64_t head_b(const MlpDesc& d, int o) {
    return head_base(d) + (d.heads2 ? (int64_t)o * (d.H + 1) + d.H : (int64_t)d.nout * d.H + o);
}
__device__ __forceinline__ uint32_t img_off(int s, int f) {  // byte offset of (sample s, feature f)
    return (uint32_t)((s >> 3) * G_S + (f >> 2) * G_F + (s & 7) * 16 + (f & 3) * 4);
}
__device__ __forceinline__ uint32_t wimg_off(int n, int k) { return (uint32_t)((n >> 3) * GW_S + (k >> 2) * G_F + (n & 7) * 16 + (k & 3) * 4); }

template <class S> __device__ void load_small_weights(S& sm, const MlpDesc& d, const float* __restrict__ p) {
    const int tid = threadIdx.x;
    const float* b1 = p + (int64_t)H * d.in;
    const float* W2 = b1 + H;
    const float* b2 = W2 + (int64_t)H * H;
    for (int k = tid; k < kInMax * H; k += NT) sm.W1[k] = (k / H) < d.in ? p[k] : 0.f;
    for (int k = tid; k < H; k += NT) { sm.b1[k] = b1[k]; sm.b2[k] = b2[k]; }
    for (int k = tid; k < H * kOutMax; k += NT) {
        int j = k / kOutMax, o = k % kOutMax;
        sm.W3[k] = o < d.nout ? p[head_w(d, o, j)] : 0.f;
    }
    if (tid < kOutMax) sm.b3[tid] = tid < d.nout ? p[head_b(d, tid)] : 0.f;
    for (int k = tid; k < H * H; k += NT) {  // W2[o + H*i]: B operand of H2pre[s][o] = sum_i H1[s][i] W2[o][i]
        int o = k % H, i = k / H;
        float w = W2[k], wh = hi_part(w);
        *reinterpret_cast<float*>(sm.B_full + wimg_off(o, i)) = w;
        *reinterpret_cast<float*>(sm.B_lo + wimg_off(o, i)) = w - wh;
    }
}

// one elected thread: D[128 x 64] (+)= A x B^T as hi*hi + hi*lo + lo*hi, K = 64 in 8 steps of 8
__device__ __forceinline__ void issue_gemm_3x(uint32_t d_tmem, const uint8_t* a_full, const uint8_t* a_lo, uint32_t a_lbo, uint32_t a_sbo,
                                              uint32_t a_kadv, const uint8_t* b_full, const uint8_t* b_lo, uint32_t b_lbo, uint32_t b_sbo,
                                              uint32_t b_kadv, uint32_t idesc, int ksteps, bool accumulate_first) {
    const uint32_t af = umma::smem_u32(a_full), al = umma::smem_u32(a_lo), bf = umma::smem_u32(b_full), bl = umma::smem_u32(b_lo);
    uint32_t acc = accumulate_first ? 1u : 0u;
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = pass == 2 ? al : af;
        const uint32_t b = pass == 1 ? bl : bf;
        for (int k = 0; k < ksteps; ++k) {
            umma::mma_tf32(d_tmem, umma::make_desc(a + k * a_kadv, a_lbo, a_sbo), umma::make_desc(b + k * b_kadv, b_lbo, b_sbo), idesc, acc);
            acc = 1u;
        }
    }
}

__device__ __forceinline__ void load_rng32(const unsigned long long* rng, int64_t i, unsigned long long (&s)[4]) {
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(rng + 4 * i);
    ulonglong2 a = p[0], b = p[1];
    s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y;
}
__device__ __forceinline__ void store_rng32(unsigned long long* rng, int64_t i, const unsigned long long (&s)[4]) {
    ulonglong2* p = reinterpret_cast<ulonglong2*>(rng + 4 * i);
    p[0] = make_ulonglong2(s[0], s[1]);
    p[1] = make_ulonglong2(s[2], s[3]);
}
__device__ __forceinline__ unsigned long long xo_next(unsigned long long (&s)[4]) {
    unsigned long long tmp = s[0] + s[3];
    unsigned long long res = ((tmp << 23) | (tmp >> 41)) + s[0];
    unsigned long long t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t;
    s[3] = (s[3] << 45) | (s[3] >> 19);
    return res;
}
__device__ __forceinline__ double xo_f64(unsigned long long (&s)[4]) { return (double)(xo_next(s) >> 11) * 0x1p-53; }
__device__ __forceinline__ float xo_f32(unsigned long long (&s)[4]) { return (float)((unsigned)(xo_next(s) >> 32) >> 8) * 0x1p-24f; }

// Thread <-> data mapping shared by layer 1, the TMEM epilogue and the image stores:
//   warp w: TMEM lane quadrant q = w % 4 (samples 32q .. 32q+31), column half c = w / 4 (features 32c .. 32c+31)
//   thread: sample s = 32q + lane, 32 features.
// mode 0: actor-critic rollout (CTA role = blockIdx & 1), mode 1: plain forward of `actor` -> head_out
__global__ void __launch_bounds__(NT, 2)
forward_tc_kernel(MlpDesc actor, MlpDesc critic, const float* __restrict__ params, AcHyper hp, int mode, const float* __restrict__ obs,
                  int64_t N, unsigned long long* __restrict__ rng, void* __restrict__ action_out, float* __restrict__ logp_out,
                  float* __restrict__ value_out, float* __restrict__ head_out, float* __restrict__ state_copy) {
    extern __shared__ unsigned char smem_raw[];
    SmemFwd& sm = *reinterpret_cast<SmemFwd*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const int nroles = mode == 0 ? 2 : 1;
    const int role = mode == 0 ? (blockIdx.x & 1) : 0;
    const int cta = blockIdx.x / nroles, nctas = gridDim.x / nroles;
    const MlpDesc d = role ? critic : actor;
    const int64_t poff = role ? actor.nparams() : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, c = warp >> 2;
    const int s = 32 * q + lane;
    load_small_weights(sm, d, params + poff);
    if (warp == 0) umma::tmem_alloc(&sm.tmem, 64);
    if (tid == 32) umma::mbar_init(&sm.bar, 1);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = sm.tmem;
    const uint32_t idesc = umma::make_idesc_tf32(128, 64, 0, 0);
    const int64_t ntiles = (N + TM - 1) / TM;
    uint32_t phase = 0;
    for (int64_t tile = cta; tile < ntiles; tile += nctas) {
        if (tid < TM) {
            int64_t i = tile * TM + tid;
            float x[kInMax] = {0.f, 0.f, 0.f, 0.f};
            if (i < N) {
                if (d.in == 4) {
                    float4 v4 = reinterpret_cast<const float4*>(obs)[i];
                    x[0] = v4.x; x[1] = v4.y; x[2] = v4.z; x[3] = v4.w;
                    if (state_copy && role == 0) reinterpret_cast<float4*>(state_copy)[i] = v4;
                } else {
                    for (int k = 0; k < d.in; ++k) {
                        x[k] = obs[(int64_t)d.in * i + k];
                        if (state_copy && role == 0) state_copy[(int64_t)d.in * i + k] = x[k];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kInMax; ++k) sm.X[k * TM + tid] = x[k];
        }
        __syncthreads();
        {   // layer 1 -> H1 operand images (full + lo)
            float x[kInMax];
#pragma unroll
            for (int k = 0; k < kInMax; ++k) x[k] = sm.X[k * TM + s];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const int f0 = 32 * c + 4 * ch;
                float4 bb = *reinterpret_cast<const float4*>(sm.b1 + f0);
                float h[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int k = 0; k < kInMax; ++k) {
                    float4 w = *reinterpret_cast<const float4*>(sm.W1 + k * H + f0);
                    h[0] = fmaf(w.x, x[k], h[0]); h[1] = fmaf(w.y, x[k], h[1]); h[2] = fmaf(w.z, x[k], h[2]); h[3] = fmaf(w.w, x[k], h[3]);
                }
                float l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) { h[e] = act_f(d.act, h[e]); l[e] = h[e] - hi_part(h[e]); }
                const uint32_t off = img_off(s, f0);
                *reinterpret_cast<float4*>(sm.A_full + off) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(sm.A_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
            }
        }
        umma::fence_proxy_async();     // generic-proxy stores -> visible to the tensor core
        umma::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            issue_gemm_3x(tmem, sm.A_full, sm.A_lo, G_F, G_S, 2 * G_F, sm.B_full, sm.B_lo, G_F, GW_S, 2 * G_F, idesc, 8, false);
            umma::commit(&sm.bar);
        }
        umma::mbar_wait(&sm.bar, phase);
        phase ^= 1u;
        umma::fence_after_sync();
        {   // epilogue: H2 = act(D + b2); head partial over this thread's 32 features
            float v[32];
            umma::tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + 32 * c, v);
            float zp[kOutMax] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int f = 32 * c + k;
                float h2 = act_f(d.act, v[k] + sm.b2[f]);
                float4 w = *reinterpret_cast<const float4*>(sm.W3 + f * kOutMax);
                zp[0] = fmaf(w.x, h2, zp[0]); zp[1] = fmaf(w.y, h2, zp[1]); zp[2] = fmaf(w.z, h2, zp[2]); zp[3] = fmaf(w.w, h2, zp[3]);
            }
#pragma unroll
            for (int o = 0; o < kOutMax; ++o) sm.Zp[(c * kOutMax + o) * TM + s] = zp[o];
        }
        umma::fence_before_sync();     // TMEM reads done before the next tile's MMA overwrites D
        __syncthreads();
        if (tid < TM) {
            int64_t i = tile * TM + tid;
            if (i < N) {
                float z[kOutMax];
#pragma unroll
                for (int o = 0; o < kOutMax; ++o) z[o] = sm.b3[o] + sm.Zp[o * TM + tid] + sm.Zp[(kOutMax + o) * TM + tid];
                if (head_out && (mode == 1 || role == 0))
                    for (int o = 0; o < d.nout; ++o) head_out[(int64_t)d.nout * i + o] = z[o];
                if (mode == 0 && role == 1) {
                    if (value_out) value_out[i] = z[0];
                } else if (mode == 0) {
                    unsigned long long st[4];
                    load_rng32(rng, i, st);
                    if (!actor.heads2) {  // sample_categorical (networks.jl:425-432)
                        int na = actor.nout;
                        float lp[kOutMax];
                        float m = -3.4e38f;
#pragma unroll
                        for (int o = 0; o < kOutMax; ++o) if (o < na) m = fmaxf(m, z[o]);
                        float se = 0.f;
#pragma unroll
                        for (int o = 0; o < kOutMax; ++o) if (o < na) se += expf(z[o] - m);
                        float ls = logf(se);
                        int best = 0;
                        double bv = 0.0;
                        float blp = 0.f;
#pragma unroll
                        for (int o = 0; o < kOutMax; ++o) {
                            if (o < na) {
                                lp[o] = (z[o] - m) - ls;
                                double u = xo_f64(st);
                                double gv = -log(-log(u)) + (double)lp[o];
                                if (o == 0 || gv > bv) { bv = gv; best = o; blp = lp[o]; }
                            }
                        }
                        if (action_out) reinterpret_cast<int32_t*>(action_out)[i] = best + 1;
                        if (logp_out) logp_out[i] = blp;
                    } else {  // GaussianNetwork
                        float mu = z[0], raw = z[1];
                        float sigma = fminf(fmaxf(softplus_f(raw), hp.min_sigma), hp.max_sigma);
                        float u1 = xo_f32(st), u2 = xo_f32(st);
                        float n = sqrtf(-2.0f * logf(1.0f - u1)) * cosf(6.2831855f * u2);
                        float a = mu + sigma * n;
                        if (action_out) reinterpret_cast<float*>(action_out)[i] = a;
                        if (logp_out) logp_out[i] = normlogpdf1(mu, sigma, a);
                    }
                    store_rng32(rng, i, st);
                }
            }
        }
        __syncthreads();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

}  // namespace

bool nn_tc_supported(const MlpDesc& d) { return d.H == 64 && d.in <= kInMax && d.nout <= kOutMax; }

int nn_tc_forward(b200rl_ctx* ctx, int grid, const MlpDesc& actor, const MlpDesc& critic, const float* params, const AcHyper& hp, int mode,
                  const float* obs, int64_t N, unsigned long long* rng, void* action_out, float* logp_out, float* value_out, float* head_out,
                  float* state_copy) {
    size_t smem = sizeof(SmemFwd) + 128;
    CUDA_TRY(cudaFuncSetAttribute(forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    forward_tc_kernel<<<grid, NT, smem, ctx->stream>>>(actor, critic, params, hp, mode, obs, N, rng, action_out, logp_out, value_out, head_out,
                                                       state_copy);
    LAUNCH_CHECK(ctx);
    return B200RL_OK;
}
