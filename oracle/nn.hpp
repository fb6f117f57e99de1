// ORACLE — TEST INFRASTRUCTURE ONLY (see jl_math.hpp header).
//
// CPU restatement of the learner side of the hot path.
// In-tree pieces followed line by line:
//   clip_by_global_norm!            RLCore/src/utils/basic.jl:19-29
//   sample_categorical (Gumbel-max) RLCore/src/utils/networks.jl:405-432 (Float64 uniforms)
//   GaussianNetwork / diagnormlogpdf RLCore/src/utils/networks.jl:44-116, distributions.jl:9-34
//   TargetNetwork sync              RLCore/src/policies/learners/target_network.jl:70-88
// Out-of-tree pieces (ReinforcementLearningZoo PPO/A2C/DQN, Optimisers.jl Adam, Flux Dense,
// all absent from /root/reference and un-pinned: no Manifest) are restated from their
// published definitions as recorded in SURVEY Appendix B — PARITY UNPINNED for those; the
// tests additionally cross-check gradients against PyTorch autograd and Adam against
// torch.optim.Adam (tests/test_oracle_nn.py).
//
// Network = Dense(in,H,act) -> Dense(H,H,act) -> heads, each head a Dense(H,d_k); parameters
// flat in Flux.destructure order, weights (out,in) column-major:
//   W1[o + H*i], b1[H], W2[o + H*i], b2[H], then per head: Wk[o + d_k*i], bk[d_k].
// Gradients are accumulated in double (this is the reference answer the fp32 GPU kernels are
// compared with to 1e-5 relative).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "jl_rng.hpp"

namespace oracle {

enum Act { ACT_RELU = 0, ACT_TANH = 1 };
constexpr int kMaxH = 256;  // per-sample scratch lives on the stack (no malloc in the hot loops)

struct MLP {
    int in, H, act, nheads;
    int hd[2];
    int out() const { return hd[0] + (nheads > 1 ? hd[1] : 0); }
    int64_t nparams() const { return (int64_t)H * in + H + (int64_t)H * H + H + (int64_t)out() * H + out(); }
    int64_t oW1() const { return 0; }
    int64_t ob1() const { return (int64_t)H * in; }
    int64_t oW2() const { return ob1() + H; }
    int64_t ob2() const { return oW2() + (int64_t)H * H; }
    int64_t oHead(int k) const { return ob2() + H + (k == 0 ? 0 : (int64_t)hd[0] * H + hd[0]); }
};

static inline float actf(int act, float z) { return act == ACT_RELU ? (z > 0 ? z : 0.f) : std::tanh(z); }
static inline float dact_from_out(int act, float h) { return act == ACT_RELU ? (h > 0 ? 1.f : 0.f) : 1.f - h * h; }

// forward: h1, h2 (size H each), out (size net.out())
static inline void mlp_forward(const MLP& n, const float* p, const float* x, float* h1, float* h2, float* out) {
    const float *W1 = p + n.oW1(), *b1 = p + n.ob1(), *W2 = p + n.oW2(), *b2 = p + n.ob2();
    for (int o = 0; o < n.H; ++o) {
        float z = b1[o];
        for (int i = 0; i < n.in; ++i) z += W1[o + n.H * i] * x[i];
        h1[o] = actf(n.act, z);
    }
    for (int o = 0; o < n.H; ++o) {
        float z = b2[o];
        for (int i = 0; i < n.H; ++i) z += W2[o + n.H * i] * h1[i];
        h2[o] = actf(n.act, z);
    }
    int off = 0;
    for (int k = 0; k < n.nheads; ++k) {
        const float* W = p + n.oHead(k);
        const float* b = W + (int64_t)n.hd[k] * n.H;
        for (int o = 0; o < n.hd[k]; ++o) {
            float z = b[o];
            for (int i = 0; i < n.H; ++i) z += W[o + n.hd[k] * i] * h2[i];
            out[off + o] = z;
        }
        off += n.hd[k];
    }
}
// backward: accumulate d(loss)/d(params) into g (double), given dout
static inline void mlp_backward(const MLP& n, const float* p, const float* x, const float* h1, const float* h2, const float* dout,
                                double* g) {
    float dh2[kMaxH], dh1[kMaxH];
    for (int i = 0; i < n.H; ++i) { dh2[i] = 0.f; dh1[i] = 0.f; }
    int off = 0;
    for (int k = 0; k < n.nheads; ++k) {
        const float* W = p + n.oHead(k);
        double* gW = g + n.oHead(k);
        double* gb = gW + (int64_t)n.hd[k] * n.H;
        for (int o = 0; o < n.hd[k]; ++o) {
            float d = dout[off + o];
            gb[o] += d;
            for (int i = 0; i < n.H; ++i) {
                gW[o + n.hd[k] * i] += (double)d * h2[i];
                dh2[i] += W[o + n.hd[k] * i] * d;
            }
        }
        off += n.hd[k];
    }
    const float* W2 = p + n.oW2();
    double *gW2 = g + n.oW2(), *gb2 = g + n.ob2(), *gW1 = g + n.oW1(), *gb1 = g + n.ob1();
    for (int o = 0; o < n.H; ++o) {
        float d = dh2[o] * dact_from_out(n.act, h2[o]);
        if (d == 0.f) continue;
        gb2[o] += d;
        for (int i = 0; i < n.H; ++i) {
            gW2[o + n.H * i] += (double)d * h1[i];
            dh1[i] += W2[o + n.H * i] * d;
        }
    }
    for (int o = 0; o < n.H; ++o) {
        float d = dh1[o] * dact_from_out(n.act, h1[o]);
        if (d == 0.f) continue;
        gb1[o] += d;
        for (int i = 0; i < n.in; ++i) gW1[o + n.H * i] += (double)d * x[i];
    }
}

// ---- heads ---------------------------------------------------------------------------------
static const float LOG2PI_F = 1.8378770664093453f;  // log(2f0*pi) = Float32(log(6.2831855f0)) (distributions.jl:9)

static inline void logsoftmax(const float* z, int n, float* lp) {
    float m = z[0];
    for (int i = 1; i < n; ++i) m = std::max(m, z[i]);
    float s = 0;
    for (int i = 0; i < n; ++i) s += std::exp(z[i] - m);
    float ls = std::log(s);
    for (int i = 0; i < n; ++i) lp[i] = (z[i] - m) - ls;
}
// sample_categorical: argmax(-log(-log(u)) + logp), u = rand(rng) Float64 per logit (networks.jl:425-432)
static inline int gumbel_argmax(jl::Xoshiro& g, const float* lp, int n, double* margin) {
    int best = 0;
    double bv = 0, second = -1e300;
    for (int i = 0; i < n; ++i) {
        double u = jl::rand_f64(g);
        double v = -std::log(-std::log(u)) + (double)lp[i];
        if (i == 0) bv = v;
        else if (v > bv) { second = bv; bv = v; best = i; }
        else if (v > second) second = v;
    }
    if (margin) *margin = bv - second;
    return best;
}
static inline float softplus(float x) { return x > 0 ? x + std::log1p(std::exp(-x)) : std::log1p(std::exp(x)); }
static inline float sigmoidf(float x) { return 1.f / (1.f + std::exp(-x)); }
// diagnormlogpdf for d = 1 (distributions.jl:31-34): -0.5*(log((s+e)^2) + (x-mu)^2/(s+e)^2 + log2pi)
static inline float normlogpdf1(float mu, float sigma, float x) {
    float s = sigma + 1e-8f;
    float v = s * s;
    float d = x - mu;
    return -0.5f * ((std::log(v) + (d * d) / v) + LOG2PI_F);
}
// normlogpdf(mu, sigma, x; eps = 1f-8) (distributions.jl:18-21): z = (x - mu) / (sigma + eps); -(z^2 + log2pi) / 2 - log(sigma + eps)
static inline float normlogpdf(float mu, float sigma, float x) {
    float z = (x - mu) / (sigma + 1e-8f);
    return -(z * z + LOG2PI_F) / 2.0f - std::log(sigma + 1e-8f);
}
// diagnormlogpdf(mu, sigma, x; eps = 1f-8) over a d-vector (distributions.jl:31-34):
// v = (sigma + eps)^2; -0.5 * (log(prod(v)) + sum((x - mu)^2 / v) + d * log2pi)
static inline float diagnormlogpdf(const float* mu, const float* sigma, const float* x, int d) {
    float prod = 1.f, sum = 0.f;
    for (int i = 0; i < d; ++i) {
        float s = sigma[i] + 1e-8f, v = s * s, e = x[i] - mu[i];
        prod *= v;
        sum += (e * e) / v;
    }
    return -0.5f * ((std::log(prod) + sum) + (float)d * LOG2PI_F);
}
// Standard normal from two Float32 uniforms of the policy stream (Box–Muller; the reference's
// randn ziggurat is not restated — SURVEY Appendix A.4 "host-independent noise definition").
static inline float randn_boxmuller(jl::Xoshiro& g) {
    float u1 = jl::rand_f32(g), u2 = jl::rand_f32(g);
    float r = std::sqrt(-2.0f * std::log(1.0f - u1));
    return r * std::cos(6.2831855f * u2);
}

// ---- actor-critic descriptors ---------------------------------------------------------------
struct ActorCritic {
    MLP actor, critic;
    int64_t nparams() const { return actor.nparams() + critic.nparams(); }
};

struct Hyper {
    float gamma, lambda, clip_range, max_grad_norm, w_actor, w_critic, w_entropy;
    float lr, beta1, beta2, eps;
    float min_sigma, max_sigma;
    int normalize_adv;
};

// policy inference for one sample (rollout): discrete -> action (1-based), logp, value
static inline void act_discrete(const ActorCritic& ac, const float* p, const float* x, jl::Xoshiro& g, int32_t* action, float* logp,
                                float* value, float* logits_out, double* margin) {
    float h1s[kMaxH], h2s[kMaxH];
    struct { float* p; float* data() { return p; } } h1{h1s}, h2{h2s};
    float z[8], lp[8], v;
    mlp_forward(ac.actor, p, x, h1.data(), h2.data(), z);
    int na = ac.actor.hd[0];
    logsoftmax(z, na, lp);
    int a = gumbel_argmax(g, lp, na, margin);
    *action = a + 1;
    *logp = lp[a];
    mlp_forward(ac.critic, p + ac.actor.nparams(), x, h1.data(), h2.data(), &v);
    *value = v;
    if (logits_out) for (int i = 0; i < na; ++i) logits_out[i] = z[i];
}
// Gaussian (1-d action): action = mu + sigma*n (unclamped; the env clamps the torque)
static inline void act_gaussian(const ActorCritic& ac, const Hyper& hp, const float* p, const float* x, jl::Xoshiro& g, float* action,
                                float* logp, float* value, float* mu_sigma_out) {
    float h1s[kMaxH], h2s[kMaxH];
    struct { float* p; float* data() { return p; } } h1{h1s}, h2{h2s};
    float z[2], v;
    mlp_forward(ac.actor, p, x, h1.data(), h2.data(), z);
    float mu = z[0];
    float sigma = std::min(std::max(softplus(z[1]), hp.min_sigma), hp.max_sigma);
    float n = randn_boxmuller(g);
    float a = mu + sigma * n;
    *action = a;
    *logp = normlogpdf1(mu, sigma, a);
    mlp_forward(ac.critic, p + ac.actor.nparams(), x, h1.data(), h2.data(), &v);
    *value = v;
    if (mu_sigma_out) { mu_sigma_out[0] = mu; mu_sigma_out[1] = sigma; }
}

// ---- losses + gradients over a minibatch ----------------------------------------------------
// Sample j of the minibatch is flat index idx[j] (or j when idx == null) into the rollout
// arrays: states (ns, total) column-major, actions/logp_old/adv/ret (total).
// losses: [actor_loss, critic_loss, entropy, loss]
struct Batch {
    const float* states; int ns;
    const int32_t* actions_i; const float* actions_f;
    const float* logp_old; const float* adv; const float* ret;
    const int32_t* idx; int64_t B;
    float adv_mean, adv_inv_std;  // (adv - mean) * inv_std applied when normalize_adv
};

template <int ALGO /*0 PPO-discrete, 1 A2C-gaussian, 2 PPO-gaussian, 3 A2C-discrete*/>
static void ac_loss_grad(const ActorCritic& ac, const Hyper& hp, const float* p, const Batch& b, double* grad, double* losses) {
    int64_t np = ac.nparams();
    std::fill(grad, grad + np, 0.0);
    double L_actor = 0, L_critic = 0, ENT = 0;
    const float invB = 1.0f / (float)b.B;
#pragma omp parallel
    {
        std::vector<double> g(np, 0.0);
        std::vector<float> h1a(ac.actor.H), h2a(ac.actor.H), h1c(ac.critic.H), h2c(ac.critic.H);
        double la = 0, lc = 0, en = 0;
#pragma omp for schedule(static)
        for (int64_t j = 0; j < b.B; ++j) {
            int64_t s = b.idx ? b.idx[j] : j;
            const float* x = b.states + (int64_t)b.ns * s;
            float A = b.adv[s];
            if (hp.normalize_adv) A = (A - b.adv_mean) * b.adv_inv_std;
            float z[8], dz[8], v;
            mlp_forward(ac.actor, p, x, h1a.data(), h2a.data(), z);
            mlp_forward(ac.critic, p + ac.actor.nparams(), x, h1c.data(), h2c.data(), &v);
            float dlogp;  // d(actor loss term)/d(logp_a) before the w_actor/B factor is folded in
            float logp_a;
            if (ALGO == 0 || ALGO == 3) {
                int na = ac.actor.hd[0];
                float lp[8] = {0.f}, pr[8] = {0.f};
                logsoftmax(z, na, lp);
                float H = 0;
                for (int i = 0; i < na; ++i) { pr[i] = std::exp(lp[i]); H -= pr[i] * lp[i]; }
                int a = b.actions_i[s] - 1;
                logp_a = lp[a];
                en += H;
                float gsel;
                if (ALGO == 0) {
                    float ratio = std::exp(logp_a - b.logp_old[s]);
                    float u = ratio * A;
                    float rc = std::min(std::max(ratio, 1.0f - hp.clip_range), 1.0f + hp.clip_range);
                    float c = rc * A;
                    la += -(double)std::min(u, c);
                    bool inside = ratio >= 1.0f - hp.clip_range && ratio <= 1.0f + hp.clip_range;
                    gsel = (u < c || inside) ? u : 0.f;  // d min(u,c) / d logp_a
                } else {
                    la += -(double)(logp_a * A);
                    gsel = A;
                }
                dlogp = -hp.w_actor * invB * gsel;
                for (int i = 0; i < na; ++i) {
                    float d = dlogp * ((i == a ? 1.f : 0.f) - pr[i]);
                    d += hp.w_entropy * invB * pr[i] * (lp[i] + H);  // -w_e * dH/dz_i
                    dz[i] = d;
                }
            } else {
                float mu = z[0], raw = z[1];
                float sp = softplus(raw);
                float sigma = std::min(std::max(sp, hp.min_sigma), hp.max_sigma);
                bool clamped = sp < hp.min_sigma || sp > hp.max_sigma;
                float a = b.actions_f[s];
                logp_a = normlogpdf1(mu, sigma, a);
                float H = std::log(sigma) + 0.5f * (LOG2PI_F + 1.0f);
                en += H;
                float gsel;
                if (ALGO == 2) {
                    float ratio = std::exp(logp_a - b.logp_old[s]);
                    float u = ratio * A;
                    float rc = std::min(std::max(ratio, 1.0f - hp.clip_range), 1.0f + hp.clip_range);
                    float c = rc * A;
                    la += -(double)std::min(u, c);
                    bool inside = ratio >= 1.0f - hp.clip_range && ratio <= 1.0f + hp.clip_range;
                    gsel = (u < c || inside) ? u : 0.f;
                } else {
                    la += -(double)(logp_a * A);
                    gsel = A;
                }
                dlogp = -hp.w_actor * invB * gsel;
                float sg = sigma + 1e-8f, d = a - mu;
                float dmu = d / (sg * sg);
                float dsig = -1.0f / sg + (d * d) / (sg * sg * sg);
                float dH_dsig = 1.0f / sigma;
                dz[0] = dlogp * dmu;
                float dsigma_total = dlogp * dsig - hp.w_entropy * invB * dH_dsig;
                dz[1] = clamped ? 0.f : dsigma_total * sigmoidf(raw);
            }
            float err = b.ret[s] - v;
            lc += (double)err * err;
            float dv = -2.0f * hp.w_critic * invB * err;
            mlp_backward(ac.actor, p, x, h1a.data(), h2a.data(), dz, g.data());
            mlp_backward(ac.critic, p + ac.actor.nparams(), x, h1c.data(), h2c.data(), &dv, g.data() + ac.actor.nparams());
        }
#pragma omp critical
        {
            for (int64_t k = 0; k < np; ++k) grad[k] += g[k];
            L_actor += la; L_critic += lc; ENT += en;
        }
    }
    losses[0] = L_actor / (double)b.B;
    losses[1] = L_critic / (double)b.B;
    losses[2] = ENT / (double)b.B;
    losses[3] = hp.w_actor * losses[0] + hp.w_critic * losses[1] - hp.w_entropy * losses[2];
}

// clip_by_global_norm! (basic.jl:19-29) on a flat gradient; returns the pre-clip norm
static inline float clip_by_global_norm(float* g, int64_t n, float clip_norm) {
    double s = 0;
    for (int64_t i = 0; i < n; ++i) s += (double)g[i] * g[i];
    float gn = (float)std::sqrt(s);
    if (clip_norm <= gn) {
        float sc = clip_norm / std::max(clip_norm, gn);
        for (int64_t i = 0; i < n; ++i) g[i] *= sc;
    }
    return gn;
}
// Optimisers.jl Adam apply!; beta_t = {beta1^t, beta2^t} carried in state
static inline void adam_step(float* p, const float* g, float* m, float* v, float* beta_t, int64_t n, float lr, float b1, float b2,
                             float eps) {
    for (int64_t i = 0; i < n; ++i) {
        m[i] = b1 * m[i] + (1 - b1) * g[i];
        v[i] = b2 * v[i] + (1 - b2) * (g[i] * g[i]);
        float dx = m[i] / (1 - beta_t[0]) / (std::sqrt(v[i] / (1 - beta_t[1])) + eps) * lr;
        p[i] -= dx;
    }
    beta_t[0] *= b1;
    beta_t[1] *= b2;
}
// TargetNetwork sync (target_network.jl:70-88): target = rho*target + (1-rho)*model
static inline void target_sync(float* target, const float* model, int64_t n, float rho) {
    for (int64_t i = 0; i < n; ++i) target[i] = rho * target[i] + (1 - rho) * model[i];
}

// ---- DQN ------------------------------------------------------------------------------------
// batch: s (ns,B), a (B, 1-based), r, t, s' (ns,B), w (B) importance weights (null = 1)
// loss = sum_i w_i * l(R_i - q_i) / B; outputs td (B) = R - q
static void dqn_loss_grad(const MLP& q, const float* p, const float* p_target, int ns, const float* s, const int32_t* a, const float* r,
                          const uint8_t* t, const float* s2, const float* w, int64_t B, float gamma, int huber, int double_dqn,
                          double* grad, double* loss_out, float* td_out) {
    int64_t np = q.nparams();
    std::fill(grad, grad + np, 0.0);
    double LOSS = 0;
    const float invB = 1.0f / (float)B;
#pragma omp parallel
    {
        std::vector<double> g(np, 0.0);
        std::vector<float> h1(q.H), h2(q.H);
        double ls = 0;
#pragma omp for schedule(static)
        for (int64_t j = 0; j < B; ++j) {
            float qn[8], qo[8], qv[8], dz[8];
            int na = q.hd[0];
            mlp_forward(q, p_target, s2 + (int64_t)ns * j, h1.data(), h2.data(), qn);
            float qnext;
            if (double_dqn) {
                mlp_forward(q, p, s2 + (int64_t)ns * j, h1.data(), h2.data(), qo);
                int best = 0;
                for (int i = 1; i < na; ++i) if (qo[i] > qo[best]) best = i;
                qnext = qn[best];
            } else {
                qnext = qn[0];
                for (int i = 1; i < na; ++i) qnext = std::max(qnext, qn[i]);
            }
            float R = r[j] + gamma * (t[j] ? 0.f : 1.f) * qnext;
            mlp_forward(q, p, s + (int64_t)ns * j, h1.data(), h2.data(), qv);
            int ai = a[j] - 1;
            float e = R - qv[ai];
            td_out[j] = e;
            float wi = w ? w[j] : 1.f;
            float ae = std::fabs(e);
            float l, dl;  // dl = d l / d q
            if (huber) {
                if (ae < 1.0f) { l = 0.5f * e * e; dl = -e; }
                else { l = ae - 0.5f; dl = e > 0 ? -1.f : 1.f; }
            } else { l = e * e; dl = -2.0f * e; }
            ls += (double)wi * l;
            for (int i = 0; i < na; ++i) dz[i] = 0.f;
            dz[ai] = wi * invB * dl;
            mlp_backward(q, p, s + (int64_t)ns * j, h1.data(), h2.data(), dz, g.data());
        }
#pragma omp critical
        {
            for (int64_t k = 0; k < np; ++k) grad[k] += g[k];
            LOSS += ls;
        }
    }
    *loss_out = LOSS / (double)B;
}

}  // namespace oracle
