// traj.cu — K3/K4: device-resident trajectory ring (CircularArraySARTSTraces), push!, and
// uniform / prioritised minibatch sampling + gather.
//
// Upstream = ReinforcementLearningTrajectories 0.4 (not vendored); anchored
// on the reference's call sites: push order agent_base.jl:45-59 / agent_srt_cache.jl:30-50,
// trace layout docs/src/How_to_implement_a_new_algorithm.md:84-112, length semantics
// RLCore/test/policies/agent.jl:27-34, iteration tuple test/policies/q_based_policy.jl:40-58.
//
// Layout in HBM: one ring of cap+1 frames; a frame holds all `lanes` sub-envs:
//   state (ns, lanes, cap+1) f32 | action (lanes, cap+1) i32 | reward (lanes, cap+1) f32 |
//   terminal (lanes, cap+1) u8 | sum tree 2L f32 (leaf = physical slot frame*lanes + lane).
// Transition j uses state frame j and (as :next_state) frame j+1 — MultiplexTraces without a
// second copy.  A push writes one contiguous frame (coalesced); a sample is one thread per
// batch slot: Xoshiro draw -> (sum-tree descent) -> 2 x state gather + scalars.
#include "common.cuh"

namespace {

struct Xo4 { unsigned long long s[4]; };
__device__ __forceinline__ Xo4 load_xo(const unsigned long long* rng, int64_t i) {
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(rng + 4 * i);
    ulonglong2 a = p[0], b = p[1];
    Xo4 g; g.s[0] = a.x; g.s[1] = a.y; g.s[2] = b.x; g.s[3] = b.y;
    return g;
}
__device__ __forceinline__ void store_xo(unsigned long long* rng, int64_t i, const Xo4& g) {
    ulonglong2* p = reinterpret_cast<ulonglong2*>(rng + 4 * i);
    p[0] = make_ulonglong2(g.s[0], g.s[1]);
    p[1] = make_ulonglong2(g.s[2], g.s[3]);
}
__device__ __forceinline__ unsigned long long xo_next(Xo4& g) {
    unsigned long long tmp = g.s[0] + g.s[3];
    unsigned long long res = ((tmp << 23) | (tmp >> 41)) + g.s[0];
    unsigned long long t = g.s[1] << 17;
    g.s[2] ^= g.s[0]; g.s[3] ^= g.s[1]; g.s[1] ^= g.s[2]; g.s[0] ^= g.s[3]; g.s[2] ^= t;
    g.s[3] = (g.s[3] << 45) | (g.s[3] >> 19);
    return res;
}
// rand(rng, Base.OneTo(n)) - 1  (Lemire nearly-divisionless, Julia SamplerRangeNDL)
__device__ __forceinline__ unsigned long long rand_below(Xo4& g, unsigned long long n) {
    unsigned long long x = xo_next(g);
    unsigned long long hi = __umul64hi(x, n), lo = x * n;
    if (lo < n) {
        unsigned long long t = (0ull - n) % n;
        while (lo < t) { x = xo_next(g); hi = __umul64hi(x, n); lo = x * n; }
    }
    return hi;
}

struct Ring {
    int ns;
    int64_t lanes, cap, first, n_states;
    float* state; int32_t* action; float* reward; uint8_t* terminal;
    float* tree; int64_t L;
    __host__ __device__ int64_t frames() const { return cap + 1; }
    __host__ __device__ int64_t phys(int64_t j) const { return (first + j) % (cap + 1); }
};

__global__ void copy_frame_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
__global__ void push_sart_kernel(Ring r, int64_t pf, const int32_t* __restrict__ a, const float* __restrict__ rew,
                                 const uint8_t* __restrict__ term) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.lanes) return;
    r.action[pf * r.lanes + e] = a[e];
    r.reward[pf * r.lanes + e] = rew[e];
    r.terminal[pf * r.lanes + e] = term[e] & 1;
}

// Single CTA: set leaves [s0, s0+n) = v0 and [s1, s1+n) = v1, then recompute their ancestors
// level by level (children are re-added, never delta-updated -> deterministic, drift-free).
__global__ void __launch_bounds__(1024) tree_set_ranges_kernel(float* __restrict__ tree, int64_t L, int64_t s0, float v0, int64_t s1,
                                                               float v1, int64_t n, int use1) {
    for (int pass = 0; pass < (use1 ? 2 : 1); ++pass) {
        int64_t s = pass ? s1 : s0;
        float v = pass ? v1 : v0;
        for (int64_t k = threadIdx.x; k < n; k += blockDim.x) tree[L + s + k] = v;
        __syncthreads();
        int64_t lo = (L + s) >> 1, hi = (L + s + n - 1) >> 1;
        while (lo >= 1) {
            for (int64_t k = lo + threadIdx.x; k <= hi; k += blockDim.x) tree[k] = tree[2 * k] + tree[2 * k + 1];
            __syncthreads();
            if (lo == 1) break;
            lo >>= 1; hi >>= 1;
        }
    }
}
// Single CTA: tree[L + key[k]] = prio[k] for the batch, then rebuild the touched paths.
__global__ void __launch_bounds__(1024) tree_update_keys_kernel(float* __restrict__ tree, int64_t L, const int64_t* __restrict__ key,
                                                                const float* __restrict__ prio, int64_t B) {
    for (int64_t k = threadIdx.x; k < B; k += blockDim.x) tree[L + key[k]] = prio[k];
    __syncthreads();
    for (int shift = 1; (L >> shift) >= 1; ++shift) {  // one tree level per iteration, leaves' parents first
        for (int64_t k = threadIdx.x; k < B; k += blockDim.x) {
            int64_t node = (L + key[k]) >> shift;
            tree[node] = tree[2 * node] + tree[2 * node + 1];
        }
        __syncthreads();
    }
}

struct BatchOut {
    float* s; int32_t* a; float* r; uint8_t* t; float* s2; int64_t* key; float* prio; float* w;
};

template <bool PRIO>
__global__ void sample_gather_kernel(Ring r, unsigned long long* __restrict__ slots, int64_t B, float beta, BatchOut o) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= B) return;
    Xo4 g = load_xo(slots, k);
    const int64_t n = (r.n_states - 1) * r.lanes;
    int64_t key, q;
    float p = 0.f, w = 1.f;
    if (PRIO) {
        float total = r.tree[1];
        float v = ((float)((unsigned)(xo_next(g) >> 32) >> 8) * 0x1p-24f) * total;  // rand(rng, Float32) * total
        int64_t node = 1;
        while (node < r.L) {
            int64_t l = 2 * node;
            float tl = r.tree[l];
            if (v <= tl) node = l;
            else { v -= tl; node = l + 1; }
        }
        key = node - r.L;
        p = r.tree[r.L + key];
        if (!(p > 0.f)) {  // rounding landed on an empty leaf: oldest transition instead
            key = r.phys(0) * r.lanes;
            p = r.tree[r.L + key];
        }
        int64_t pf = key / r.lanes, e = key % r.lanes;
        int64_t j = (pf - r.first + r.frames()) % r.frames();
        q = j * r.lanes + e;
        w = powf((float)n * (p / total), -beta);
    } else {
        q = (int64_t)rand_below(g, (unsigned long long)n);
        key = r.phys(q / r.lanes) * r.lanes + q % r.lanes;
    }
    store_xo(slots, k, g);
    int64_t j = q / r.lanes, e = q % r.lanes;
    int64_t pf = r.phys(j), pn = r.phys(j + 1);
    const float* s = r.state + (int64_t)r.ns * (pf * r.lanes + e);
    const float* s2 = r.state + (int64_t)r.ns * (pn * r.lanes + e);
    for (int c = 0; c < r.ns; ++c) { o.s[(int64_t)r.ns * k + c] = s[c]; o.s2[(int64_t)r.ns * k + c] = s2[c]; }
    o.a[k] = r.action[pf * r.lanes + e];
    o.r[k] = r.reward[pf * r.lanes + e];
    o.t[k] = r.terminal[pf * r.lanes + e];
    o.key[k] = pf * r.lanes + e;
    o.prio[k] = p;
    o.w[k] = w;
}
// w /= max(w) (single CTA, fixed tree)
__global__ void __launch_bounds__(1024) normalize_weights_kernel(float* __restrict__ w, int64_t B) {
    __shared__ float red[32];
    float m = 0.f;
    for (int64_t k = threadIdx.x; k < B; k += blockDim.x) m = fmaxf(m, w[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) m = fmaxf(m, red[k]);
    for (int64_t k = threadIdx.x; k < B; k += blockDim.x) w[k] = w[k] / m;
}
// new priorities from TD errors: (|td| + eps)^alpha   (PrioritizedDQN)
__global__ void td_to_priority_kernel(const float* __restrict__ td, float* __restrict__ prio, int64_t B, float eps, float alpha) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < B) prio[k] = powf(fabsf(td[k]) + eps, alpha);
}

}  // namespace

struct b200rl_traj {
    b200rl_ctx* ctx;
    Ring r;
    bool prioritized;
    float default_priority;
    int64_t B;
    unsigned long long* slots;  // (4, B) sampler streams
    BatchOut batch;
    float* new_prio;            // (B) scratch for priority write-back
    void* stage;                // staging for host-side pushes
    size_t stage_bytes;
};

static int stage_in(b200rl_traj* t, const void* host, size_t bytes, size_t offset, const void** dev) {
    REQUIRE(offset + bytes <= t->stage_bytes, B200RL_ERR_INVALID, "staging overflow");
    CUDA_TRY(cudaMemcpyAsync((char*)t->stage + offset, host, bytes, cudaMemcpyHostToDevice, t->ctx->stream));
    *dev = (char*)t->stage + offset;
    return B200RL_OK;
}

static int traj_push_state_dev(b200rl_traj* t, const float* obs_dev) {
    Ring& r = t->r;
    int64_t pf;
    if (r.n_states == r.frames()) { pf = r.first; r.first = (r.first + 1) % r.frames(); }
    else { pf = r.phys(r.n_states); r.n_states += 1; }
    int64_t n = (int64_t)r.ns * r.lanes;
    copy_frame_kernel<<<grid_for(n, 256), 256, 0, t->ctx->stream>>>(r.state + n * pf, obs_dev, n);
    LAUNCH_CHECK(t->ctx);
    return (int)0;
}

extern "C" {

int b200rl_traj_create(b200rl_ctx* ctx, int ns, int64_t lanes, int64_t capacity, int prioritized, float default_priority,
                       const uint64_t* sampler_rng, int64_t batch_size, b200rl_traj** out) {
    TRY(ctx_bind(ctx));
    REQUIRE(out && ns >= 1 && ns <= 16 && lanes >= 1 && capacity >= 1, B200RL_ERR_INVALID, "bad shape");
    REQUIRE(batch_size >= 0 && (batch_size == 0 || sampler_rng), B200RL_ERR_INVALID, "sampler_rng required when batch_size > 0");
    b200rl_traj* t = new b200rl_traj();
    memset(t, 0, sizeof *t);
    t->ctx = ctx; t->prioritized = prioritized != 0; t->default_priority = default_priority; t->B = batch_size;
    Ring& r = t->r;
    r.ns = ns; r.lanes = lanes; r.cap = capacity; r.first = 0; r.n_states = 0;
    size_t slots = (size_t)lanes * (capacity + 1);
    CUDA_TRY(cudaMalloc(&r.state, slots * ns * sizeof(float)));
    CUDA_TRY(cudaMalloc(&r.action, slots * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&r.reward, slots * sizeof(float)));
    CUDA_TRY(cudaMalloc(&r.terminal, slots));
    r.L = 1;
    if (t->prioritized) {
        while (r.L < (int64_t)slots) r.L <<= 1;
        CUDA_TRY(cudaMalloc(&r.tree, 2 * r.L * sizeof(float)));
        CUDA_TRY(cudaMemsetAsync(r.tree, 0, 2 * r.L * sizeof(float), ctx->stream));
    }
    if (batch_size > 0) {
        size_t B = (size_t)batch_size;
        CUDA_TRY(cudaMalloc(&t->slots, B * 32));
        CUDA_TRY(cudaMemcpyAsync(t->slots, sampler_rng, B * 32, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMalloc(&t->batch.s, B * ns * sizeof(float)));
        CUDA_TRY(cudaMalloc(&t->batch.s2, B * ns * sizeof(float)));
        CUDA_TRY(cudaMalloc(&t->batch.a, B * 4));
        CUDA_TRY(cudaMalloc(&t->batch.r, B * 4));
        CUDA_TRY(cudaMalloc(&t->batch.t, B));
        CUDA_TRY(cudaMalloc(&t->batch.key, B * 8));
        CUDA_TRY(cudaMalloc(&t->batch.prio, B * 4));
        CUDA_TRY(cudaMalloc(&t->batch.w, B * 4));
        CUDA_TRY(cudaMalloc(&t->new_prio, B * 4));
    }
    t->stage_bytes = (size_t)lanes * (ns * 4 + 4 + 4 + 1) + 64;
    CUDA_TRY(cudaMalloc(&t->stage, t->stage_bytes));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *out = t;
    return B200RL_OK;
}

int b200rl_traj_destroy(b200rl_traj* t) {
    if (!t) return B200RL_OK;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->r.state); cudaFree(t->r.action); cudaFree(t->r.reward); cudaFree(t->r.terminal); cudaFree(t->r.tree);
    cudaFree(t->slots); cudaFree(t->batch.s); cudaFree(t->batch.s2); cudaFree(t->batch.a); cudaFree(t->batch.r); cudaFree(t->batch.t);
    cudaFree(t->batch.key); cudaFree(t->batch.prio); cudaFree(t->batch.w); cudaFree(t->new_prio); cudaFree(t->stage);
    delete t;
    return B200RL_OK;
}

/* length(trajectory.container): number of complete transition frames (x lanes transitions) */
int b200rl_traj_length(b200rl_traj* t, int64_t* frames_out) {
    REQUIRE(t && frames_out, B200RL_ERR_INVALID, "null argument");
    *frames_out = t->r.n_states > 0 ? t->r.n_states - 1 : 0;
    return B200RL_OK;
}

int b200rl_traj_push_state(b200rl_traj* t, const float* obs, int on_device) {
    REQUIRE(t && obs, B200RL_ERR_INVALID, "null argument");
    TRY(ctx_bind(t->ctx));
    const void* d = obs;
    if (!on_device) TRY(stage_in(t, obs, (size_t)t->r.ns * t->r.lanes * 4, 0, &d));
    // the frame being overwritten is the oldest one: its transitions leave the sum tree
    int64_t pf = t->r.n_states == t->r.frames() ? t->r.first : t->r.phys(t->r.n_states);
    TRY(traj_push_state_dev(t, (const float*)d));
    if (t->prioritized) {
        tree_set_ranges_kernel<<<1, 1024, 0, t->ctx->stream>>>(t->r.tree, t->r.L, pf * t->r.lanes, 0.f, 0, 0.f, t->r.lanes, 0);
        LAUNCH_CHECK(t->ctx);
    }
    if (!on_device) CUDA_TRY(cudaStreamSynchronize(t->ctx->stream));
    return B200RL_OK;
}

int b200rl_traj_push(b200rl_traj* t, const int32_t* action, const float* reward, const uint8_t* terminal, const float* next_obs,
                     int on_device) {
    REQUIRE(t && action && reward && terminal && next_obs, B200RL_ERR_INVALID, "null argument");
    REQUIRE(t->r.n_states >= 1, B200RL_ERR_INVALID, "push a first state (b200rl_traj_push_state) before the first transition");
    TRY(ctx_bind(t->ctx));
    Ring& r = t->r;
    const void *da = action, *dr = reward, *dt = terminal, *ds = next_obs;
    if (!on_device) {
        size_t L = (size_t)r.lanes;
        TRY(stage_in(t, next_obs, L * r.ns * 4, 0, &ds));
        TRY(stage_in(t, action, L * 4, L * r.ns * 4, &da));
        TRY(stage_in(t, reward, L * 4, L * r.ns * 4 + L * 4, &dr));
        TRY(stage_in(t, terminal, L, L * r.ns * 4 + L * 8, &dt));
    }
    int64_t pf = r.phys(r.n_states - 1);
    push_sart_kernel<<<grid_for(r.lanes, 256), 256, 0, t->ctx->stream>>>(r, pf, (const int32_t*)da, (const float*)dr, (const uint8_t*)dt);
    LAUNCH_CHECK(t->ctx);
    int64_t pnew = r.n_states == r.frames() ? r.first : r.phys(r.n_states);
    TRY(traj_push_state_dev(t, (const float*)ds));
    if (t->prioritized) {
        tree_set_ranges_kernel<<<1, 1024, 0, t->ctx->stream>>>(r.tree, r.L, pnew * r.lanes, 0.f, pf * r.lanes, t->default_priority, r.lanes, 1);
        LAUNCH_CHECK(t->ctx);
    }
    if (!on_device) CUDA_TRY(cudaStreamSynchronize(t->ctx->stream));
    return B200RL_OK;
}

/* BatchSampler / prioritised sampler + gather into the trajectory's device batch buffers */
int b200rl_traj_sample(b200rl_traj* t, float beta) {
    REQUIRE(t && t->B > 0, B200RL_ERR_INVALID, "trajectory was created without a sampler");
    REQUIRE(t->r.n_states >= 2, B200RL_ERR_INVALID, "nothing to sample yet");
    TRY(ctx_bind(t->ctx));
    if (t->prioritized) {
        sample_gather_kernel<true><<<grid_for(t->B, 128), 128, 0, t->ctx->stream>>>(t->r, t->slots, t->B, beta, t->batch);
        LAUNCH_CHECK(t->ctx);
        normalize_weights_kernel<<<1, 1024, 0, t->ctx->stream>>>(t->batch.w, t->B);
        LAUNCH_CHECK(t->ctx);
    } else {
        sample_gather_kernel<false><<<grid_for(t->B, 128), 128, 0, t->ctx->stream>>>(t->r, t->slots, t->B, beta, t->batch);
        LAUNCH_CHECK(t->ctx);
    }
    return B200RL_OK;
}

/* field: 0 state (ns,B) 1 action (B) i32 2 reward 3 terminal u8 4 next_state 5 key i64 6 priority 7 weight 8 sampler rng (4,B) */
int b200rl_traj_batch_get(b200rl_traj* t, int field, void* host_dst, size_t bytes) {
    REQUIRE(t && host_dst && t->B > 0, B200RL_ERR_INVALID, "bad argument");
    TRY(ctx_bind(t->ctx));
    size_t B = (size_t)t->B;
    const void* src = nullptr;
    size_t need = 0;
    switch (field) {
        case 0: src = t->batch.s; need = B * t->r.ns * 4; break;
        case 1: src = t->batch.a; need = B * 4; break;
        case 2: src = t->batch.r; need = B * 4; break;
        case 3: src = t->batch.t; need = B; break;
        case 4: src = t->batch.s2; need = B * t->r.ns * 4; break;
        case 5: src = t->batch.key; need = B * 8; break;
        case 6: src = t->batch.prio; need = B * 4; break;
        case 7: src = t->batch.w; need = B * 4; break;
        case 8: src = t->slots; need = B * 32; break;
        default: REQUIRE(false, B200RL_ERR_INVALID, "unknown batch field");
    }
    REQUIRE(bytes >= need, B200RL_ERR_INVALID, "destination too small");
    CUDA_TRY(cudaMemcpyAsync(host_dst, src, need, cudaMemcpyDeviceToHost, t->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(t->ctx->stream));
    return B200RL_OK;
}

/* priority write-back for the keys of the last sampled batch */
int b200rl_traj_update_priority(b200rl_traj* t, const float* prio, int on_device) {
    REQUIRE(t && prio && t->prioritized && t->B > 0, B200RL_ERR_INVALID, "bad argument");
    TRY(ctx_bind(t->ctx));
    const float* d = prio;
    if (!on_device) {
        CUDA_TRY(cudaMemcpyAsync(t->new_prio, prio, (size_t)t->B * 4, cudaMemcpyHostToDevice, t->ctx->stream));
        d = t->new_prio;
    }
    tree_update_keys_kernel<<<1, 1024, 0, t->ctx->stream>>>(t->r.tree, t->r.L, t->batch.key, d, t->B);
    LAUNCH_CHECK(t->ctx);
    if (!on_device) CUDA_TRY(cudaStreamSynchronize(t->ctx->stream));
    return B200RL_OK;
}

int b200rl_traj_total_priority(b200rl_traj* t, float* out) {
    REQUIRE(t && out && t->prioritized, B200RL_ERR_INVALID, "bad argument");
    TRY(ctx_bind(t->ctx));
    CUDA_TRY(cudaMemcpyAsync(out, t->r.tree + 1, 4, cudaMemcpyDeviceToHost, t->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(t->ctx->stream));
    return B200RL_OK;
}

}  // extern "C"

// ---- internal accessors for algo.cu -----------------------------------------------------------
struct TrajBatchView { const float* s; const int32_t* a; const float* r; const uint8_t* t; const float* s2; const float* w; int64_t B; int ns; };
TrajBatchView b200rl_traj_internal_batch(b200rl_traj* t) {
    return TrajBatchView{t->batch.s, t->batch.a, t->batch.r, t->batch.t, t->batch.s2, t->prioritized ? t->batch.w : nullptr, t->B, t->r.ns};
}
bool b200rl_traj_internal_prioritized(b200rl_traj* t) { return t->prioritized; }
b200rl_ctx* b200rl_traj_internal_ctx(b200rl_traj* t) { return t->ctx; }
int64_t b200rl_traj_internal_lanes(b200rl_traj* t) { return t->r.lanes; }
int b200rl_traj_internal_priority_from_td(b200rl_traj* t, const float* td_dev, float eps, float alpha) {
    td_to_priority_kernel<<<grid_for(t->B, 256), 256, 0, t->ctx->stream>>>(td_dev, t->new_prio, t->B, eps, alpha);
    LAUNCH_CHECK(t->ctx);
    tree_update_keys_kernel<<<1, 1024, 0, t->ctx->stream>>>(t->r.tree, t->r.L, t->batch.key, t->new_prio, t->B);
    LAUNCH_CHECK(t->ctx);
    return B200RL_OK;
}
