// tc_fwd.cuh — device building blocks of the tensor-core forward pass of one 2 x 64 MLP (K6), shared by the policy
// inference kernel and the fused rollout kernel (fwd_tc.cu).  One tile = 128 samples = UMMA M.
//
//   layer 1 (K <= 4)  : FP32 FFMA, thread = (sample, 32 features), written straight into TMEM as the A operand (hi | lo fp16 pairs)
//   layer 2 (64 x 64) : tcgen05.mma kind::f16 (K = 16), TS form (A from TMEM, B = W2 image in shared memory), the 3-term fp16 split
//                       x_hi = fp16(64 x), x_lo = fp16(64 x - x_hi) (22 mantissa bits, like 3xTF32, at half the instruction count;
//                       see nn_tc.cu for the pacing measurement and the range limits):
//                         MMA 1 (N = 128): D[0:64) = hi*hi, D[64:128) = hi*lo     (rows 0..63 | 64..127 of the B image)
//                         MMA 2 (N =  64): D[0:64) += lo*hi
//   heads             : FP32 FFMA on the epilogue registers, partial sums of the two 32-feature halves meet in shared memory
//
// Thread <-> data: warp w: TMEM lane quadrant q = w % 4 (samples 32q .. 32q+31), column half c = w / 4; thread = sample
// s = 32q + lane, features 32c .. 32c+31.  Every multiply-add is an explicit fmaf / separate op: the arithmetic does not
// depend on the contraction flags of the including translation unit.
#pragma once
#include "nn.cuh"
#include "umma.cuh"

namespace tcfwd {

constexpr int NT = 256;
constexpr int TM = 128;
constexpr int H = 64;
constexpr int G_F = 128;              // byte stride between 8-element (16-byte) chunks along K: one 8 x 16 B core matrix
constexpr int GW_S = 8 * G_F;         // weight image: stride between 8-row groups (K = 64 = 8 chunks)
constexpr int WIMG_BYTES = 16 * GW_S; // [hi (64 rows) ; lo (64 rows)] x [K = 64] fp16 weight image (SWIZZLE_NONE, K-major core matrices)
constexpr uint32_t COL_D = 0, COL_A = 128;   // TMEM columns: accumulator [hh+lh | hl], A operand [hi: 32 columns of fp16 pairs | lo: 32]
constexpr float kScale = 64.0f;       // power-of-two scale of both operands (exact; undone on the accumulator)
constexpr uint32_t TMEM_COLS = 256;
constexpr float kLog2Pi = 1.8378770664093453f;

struct NetSm {   // one network's weights in shared memory
    alignas(128) uint8_t B[WIMG_BYTES];        // 64 W2 as (n = out, k = in), K-major fp16: rows 0..63 hi, 64..127 lo (one N = 128 operand)
    float W1[kInMax * H];                      // [i][o]
    float b1[H], b2[H];
    float W3[H * kOutMax];                     // [j][o]
    float b3[kOutMax];
};

__device__ __forceinline__ float act_f(int act, float z) { return act == B200RL_ACT_RELU ? fmaxf(z, 0.f) : tanhf(z); }
// two fp32 values -> packed fp16 pairs {low half = a, high half = b}: hi parts, and the fp16 of what they miss (lo parts)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h);
    const __half2 l = __floats2half2_rn(__fsub_rn(a, back.x), __fsub_rn(b, back.y));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float softplus_f(float x) { return x > 0.f ? x + log1pf(expf(-x)) : log1pf(expf(x)); }
__device__ __forceinline__ float normlogpdf1(float mu, float sigma, float x) {   // distributions.jl:18-21, eps = 1f-8
    float s = sigma + 1e-8f, v = __fmul_rn(s, s), dd = x - mu;
    return __fmul_rn(-0.5f, (logf(v) + __fmul_rn(dd, dd) / v) + kLog2Pi);
}
__device__ __forceinline__ int64_t head_base(const MlpDesc& d) { return (int64_t)d.H * d.in + d.H + (int64_t)d.H * d.H + d.H; }
__device__ __forceinline__ int64_t head_w(const MlpDesc& d, int o, int j) {
    return head_base(d) + (d.heads2 ? (int64_t)o * (d.H + 1) + j : (int64_t)o + (int64_t)d.nout * j);
}
__device__ __forceinline__ int64_t head_b(const MlpDesc& d, int o) {
    return head_base(d) + (d.heads2 ? (int64_t)o * (d.H + 1) + d.H : (int64_t)d.nout * d.H + o);
}
__device__ __forceinline__ uint32_t wimg_off(int n, int k) { return (uint32_t)((n >> 3) * GW_S + (k >> 3) * G_F + (n & 7) * 16 + (k & 7) * 2); }

// s1: scale folded into W1 / b1 (kScale for relu trunks: relu(S z) = S relu(z) exactly for a power of two S, so layer 1 then
// produces the scaled H1 operand without a multiply per feature; 1 otherwise)
__device__ inline void load_net(NetSm& w, const MlpDesc& d, const float* __restrict__ p, float s1) {
    const int tid = threadIdx.x;
    const float* b1 = p + (int64_t)H * d.in;
    const float* W2 = b1 + H;
    const float* b2 = W2 + (int64_t)H * H;
    for (int k = tid; k < kInMax * H; k += NT) w.W1[k] = (k / H) < d.in ? __fmul_rn(p[k], s1) : 0.f;
    for (int k = tid; k < H; k += NT) { w.b1[k] = __fmul_rn(b1[k], s1); w.b2[k] = b2[k]; }
    for (int k = tid; k < H * kOutMax; k += NT) {
        int j = k / kOutMax, o = k % kOutMax;
        w.W3[k] = o < d.nout ? p[head_w(d, o, j)] : 0.f;
    }
    if (tid < kOutMax) w.b3[tid] = tid < d.nout ? p[head_b(d, tid)] : 0.f;
    for (int k = tid; k < H * H; k += NT) {   // W2[o + H*i]: B operand of H2pre[s][o] = sum_i H1[s][i] W2[o][i]
        int o = k % H, i = k / H;
        const float v = __fmul_rn(W2[k], kScale);
        const __half vh = __float2half_rn(v), vl = __float2half_rn(__fsub_rn(v, __half2float(vh)));
        *reinterpret_cast<__half*>(w.B + wimg_off(o, i)) = vh;
        *reinterpret_cast<__half*>(w.B + wimg_off(H + o, i)) = vl;
    }
}

// Xoshiro256++ on a 4-word state held in registers (policy stream of one env)
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

// layer 1 of this thread's sample: H1[32c .. 32c+32) = act(W1 x + b1) -> TMEM A operand (hi pairs at COL_A, lo pairs at COL_A + 32)
// (the loops over 16-feature halves here and 8-feature groups in head_partials are deliberately NOT unrolled: tanhf is ~40
// instructions, and with every instance inlined the rollout kernel was 290 KB of SASS whose dominant stall was instruction fetch)
// ACT: the activation as a compile-time constant (-1: read `act`); relu trunks expect load_net(.., s1 = kScale)
template <int ACT>
__device__ __forceinline__ void layer1_to_tmem(const NetSm& w, int act_rt, const float (&x)[kInMax], int c, uint32_t tmem_lane) {
    const int act = ACT >= 0 ? ACT : act_rt;
#pragma unroll(ACT == B200RL_ACT_RELU ? 2 : 1)
    for (int half = 0; half < 2; ++half) {
        uint32_t hi8[8], lo8[8];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const int f0 = 32 * c + 16 * half + 4 * ch;
            float4 bb = *reinterpret_cast<const float4*>(w.b1 + f0);
            float h[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int k = 0; k < kInMax; ++k) {
                float4 ww = *reinterpret_cast<const float4*>(w.W1 + k * H + f0);
                h[0] = fmaf(ww.x, x[k], h[0]); h[1] = fmaf(ww.y, x[k], h[1]); h[2] = fmaf(ww.z, x[k], h[2]); h[3] = fmaf(ww.w, x[k], h[3]);
            }
            if (ACT == B200RL_ACT_RELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = fmaxf(h[e], 0.f);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __fmul_rn(act_f(act, h[e]), kScale);
            }
            split2(h[0], h[1], hi8[2 * ch], lo8[2 * ch]);
            split2(h[2], h[3], hi8[2 * ch + 1], lo8[2 * ch + 1]);
        }
        umma::tmem_st8(tmem_lane + COL_A + 16 * c + 8 * half, hi8);
        umma::tmem_st8(tmem_lane + COL_A + 32 + 16 * c + 8 * half, lo8);
    }
    umma::tmem_st_wait();
}

// one elected thread: D = A x W2^T as the 3-term fp16 split, all three terms accumulated into the same 64 columns
// (12 MMAs of N = 64, K = 16 each: hi*hi, hi*lo, lo*hi)
__device__ __forceinline__ void issue_gemm(uint32_t tmem, const NetSm& w) {
    const uint32_t idesc64 = umma::make_idesc_f16(128, 64, 0, 0);
    const uint64_t dB = umma::make_desc(umma::smem_u32(w.B), G_F, GW_S);
    const uint64_t dBlo = dB + (uint64_t)((8 * GW_S) >> 4);   // rows 64..127 of the image
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)(k * (2 * G_F / 16));
        umma::mma_f16_ts(tmem + COL_D, tmem + COL_A + 8 * k, dB + adv, idesc64, k ? 1u : 0u);
        umma::mma_f16_ts(tmem + COL_D, tmem + COL_A + 8 * k, dBlo + adv, idesc64, 1u);
        umma::mma_f16_ts(tmem + COL_D, tmem + COL_A + 32 + 8 * k, dB + adv, idesc64, 1u);
    }
}

// epilogue of this thread's sample: H2[32c .. 32c+32) = act(D + b2), partial head sums over these 32 features
template <int ACT>
__device__ __forceinline__ void head_partials(const NetSm& w, int act_rt, int c, uint32_t tmem_lane, float (&zp)[kOutMax]) {
    const int act = ACT >= 0 ? ACT : act_rt;
#pragma unroll
    for (int o = 0; o < kOutMax; ++o) zp[o] = 0.f;
#pragma unroll(ACT == B200RL_ACT_RELU ? 2 : 1)
    for (int grp = 0; grp < 2; ++grp) {
        float v[16];
        umma::tmem_ld16(tmem_lane + COL_D + 32 * c + 16 * grp, v);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int f = 32 * c + 16 * grp + k;
            float h2 = act_f(act, fmaf(v[k], 1.0f / (kScale * kScale), w.b2[f]));   // operand scales undone (exact power of two)
            float4 ww = *reinterpret_cast<const float4*>(w.W3 + f * kOutMax);
            zp[0] = fmaf(ww.x, h2, zp[0]); zp[1] = fmaf(ww.y, h2, zp[1]); zp[2] = fmaf(ww.z, h2, zp[2]); zp[3] = fmaf(ww.w, h2, zp[3]);
        }
    }
}

// one Gumbel(0, 1) draw in Float64 (one out-of-line copy: the double-precision log is ~150 instructions)
__device__ __noinline__ double gumbel64(double u) { return -log(-log(u)); }

// policy head: sample an action and its log-probability from the head outputs z on the env's policy stream.
// Categorical: sample_categorical (networks.jl:425-432), Float64 Gumbel noise; Gaussian: GaussianNetwork (networks.jl:64-116).
// Returns the action as raw 32 bits (int32 1-based | float).
__device__ __forceinline__ uint32_t sample_head(const MlpDesc& actor, const AcHyper& hp, const float (&z)[kOutMax], unsigned long long (&st)[4],
                                                float& logp) {
    if (!actor.heads2) {
        const int na = actor.nout;
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
                double gv = gumbel64(u) + (double)lp[o];
                if (o == 0 || gv > bv) { bv = gv; best = o; blp = lp[o]; }
            }
        }
        logp = blp;
        return (uint32_t)(best + 1);
    }
    float mu = z[0], raw = z[1];
    float sigma = fminf(fmaxf(softplus_f(raw), hp.min_sigma), hp.max_sigma);
    float u1 = xo_f32(st), u2 = xo_f32(st);
    float n = __fmul_rn(sqrtf(__fmul_rn(-2.0f, logf(1.0f - u1))), cosf(__fmul_rn(6.2831855f, u2)));
    float a = __fadd_rn(mu, __fmul_rn(sigma, n));
    logp = normlogpdf1(mu, sigma, a);
    return __float_as_uint(a);
}

}  // namespace tcfwd
