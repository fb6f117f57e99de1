"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liboracle.so")

F_STATE, F_OBS, F_REWARD, F_TERMINAL, F_T, F_RNG, F_FLAGS, F_ACTION = 0, 1, 2, 3, 4, 5, 6, 7
KIND_CARTPOLE, KIND_PENDULUM, KIND_MOUNTAINCAR, KIND_CARTPOLE_CONT, KIND_MOUNTAINCAR_CONT, KIND_ACROBOT = 0, 1, 2, 3, 4, 5
NS = {0: 4, 1: 2, 2: 2, 3: 4, 4: 2, 5: 4}
NOBS = {0: 4, 1: 3, 2: 2, 3: 4, 4: 2, 5: 6}


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def lib():
    global _L
    try:
        return _L
    except NameError:
        pass
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    vp, i32, i64, u64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_double
    sig = {
        "orc_max_threads": (i32, []), "orc_set_threads": (None, [i32]),
        "orc_sin32": (f32, [f32]), "orc_cos32": (f32, [f32]), "orc_sin64": (f64, [f64]),
        "orc_cos64": (f64, [f64]), "orc_mod64": (f64, [f64, f64]),
        "orc_rng_next": (u64, [vp]), "orc_rng_oneto": (i64, [vp, u64]),
        "orc_seed_splitmix": (None, [u64, vp]),
        "orc_jl_shuffle_i64": (None, [vp, vp, i64]), "orc_jl_shuffle_i32": (None, [vp, vp, i64]), "orc_jl_randperm": (None, [vp, vp, i64]),
        "orc_jl_ltm52": (u64, [vp, u64, u64]),
        "orc_cartpole_default_params": (None, [i32, vp]),
        "orc_vecenv_create": (vp, [i32, i32, i64, vp, vp]), "orc_vecenv_destroy": (None, [vp]),
        "orc_vecenv_set_max_timeout": (None, [vp, i64]),
        "orc_vecenv_reset": (None, [vp, i32]), "orc_vecenv_step": (i32, [vp, vp, i32]),
        "orc_vecenv_step_random": (i32, [vp, i32, vp]), "orc_vecenv_get": (None, [vp, i32, vp]),
        "orc_vecenv_set": (None, [vp, i32, vp]), "orc_vecenv_bench_random": (f64, [vp, i32]),
        "orc_gae_f32": (None, [vp, vp, vp, vp, f32, f32, i64, i64, i32]),
        "orc_gae_f64": (None, [vp, vp, vp, vp, f64, f64, i64, i64, i32]),
        "orc_discount_f32": (None, [vp, vp, vp, vp, f32, i64, i64, i32]),
        "orc_discount_f64": (None, [vp, vp, vp, vp, f64, i64, i64, i32]),
        "orc_discount_reduced_f32": (None, [vp, vp, vp, vp, f32, i64, i64, i32]),
        "orc_discount_reduced_f64": (None, [vp, vp, vp, vp, f64, i64, i64, i32]),
        "orc_ac_nparams": (i64, [vp]), "orc_q_nparams": (i64, [vp]),
        "orc_perm_index": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint32]),
        "orc_act_discrete": (None, [vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]),
        "orc_act_gaussian": (None, [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp]),
        "orc_critic_values": (None, [vp, vp, vp, i64, vp]), "orc_q_values": (None, [vp, vp, vp, i64, vp]),
        "orc_ac_loss_grad": (None, [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, f32, f32, vp, vp]),
        "orc_clip_by_global_norm": (f32, [vp, i64, f32]),
        "orc_adam_step": (None, [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32]),
        "orc_target_sync": (None, [vp, vp, i64, f32]),
        "orc_dqn_loss_grad": (None, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, f32, i32, i32, vp, vp, vp]),
        "orc_adv_norm": (None, [vp, i64, vp]),
        "orc_traj_create": (vp, [i32, i64, i64, i32, f32]), "orc_traj_destroy": (None, [vp]),
        "orc_traj_length": (i64, [vp]), "orc_traj_push_state": (None, [vp, vp]), "orc_traj_lane_lengths": (None, [vp, vp]),
        "orc_traj_n_sampleable": (i64, [vp]), "orc_traj_push_episode_start": (None, [vp, vp, i32]),
        "orc_traj_push": (None, [vp, vp, vp, vp, vp]),
        "orc_traj_sample": (None, [vp, i32, vp, i64, f32, vp, vp, vp, vp, vp, vp, vp, vp]),
        "orc_traj_update_priority": (None, [vp, vp, vp, i64]), "orc_traj_total_priority": (f32, [vp]),
        "orc_normlogpdf": (f32, [f32, f32, f32]), "orc_normlogpdf1": (f32, [f32, f32, f32]), "orc_diagnormlogpdf": (f32, [vp, vp, vp, i32]),
        "orc_get_eps": (f64, [vp, i64]), "orc_egreedy_prob": (None, [vp, i64, vp, i32, vp]),
        "orc_egreedy_plan": (None, [vp, i64, vp, i32, i64, vp, vp]),
        "orc_ppo_iteration": (f64, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, C.c_uint32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _L = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def julia_xoshiro(seed):
    """Raw state of Julia's `Xoshiro(seed::Integer)` for 1.7 <= Julia <= 1.10 (stdlib Random, Xoshiro.jl `seed!`): the SHA-256
    digest of the seed's UInt32 words (little endian, `make_seed`) read as four little-endian UInt64.  Pinned by the vectors Julia's
    manual prints (tests/test_oracle_julia_rng.py).  Julia 1.11 changed the seeding; the C ABI therefore takes raw states."""
    import hashlib
    seed = int(seed)
    if seed < 0:
        raise ValueError("seed must be non-negative")
    words = []
    while True:
        words.append(seed & 0xFFFFFFFF)
        seed >>= 32
        if seed == 0:
            break
    digest = hashlib.sha256(b"".join(w.to_bytes(4, "little") for w in words)).digest()
    return np.frombuffer(digest, dtype="<u8").astype(np.uint64)


def jl_shuffle(state, a):
    """shuffle!(rng, a) on a raw Xoshiro state (advanced in place); a: int32 or int64 array, shuffled in place."""
    a = np.ascontiguousarray(a)
    fn = lib().orc_jl_shuffle_i32 if a.dtype == np.int32 else lib().orc_jl_shuffle_i64
    assert a.dtype in (np.int32, np.int64)
    fn(_p(state), _p(a), a.size)
    return a


def jl_randperm(state, n):
    a = np.zeros(n, np.int64)
    lib().orc_jl_randperm(_p(state), _p(a), n)
    return a


def splitmix_states(n, seed=0x9E3779B97F4A7C15):
    """Test-harness seeding of SURVEY §8d: env i gets 4 splitmix64 outputs of seed ^ i."""
    out = np.empty((n, 4), dtype=np.uint64)
    tmp = np.empty(4, dtype=np.uint64)
    L = lib()
    for i in range(n):
        L.orc_seed_splitmix(C.c_uint64((seed ^ i) & 0xFFFFFFFFFFFFFFFF), _p(tmp))
        out[i] = tmp
    return out


def splitmix_states_fast(n, seed=0x9E3779B97F4A7C15):
    """Vectorised numpy version of splitmix_states (same values)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    x = (np.uint64(seed) ^ np.arange(n, dtype=np.uint64))
    out = np.empty((n, 4), dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k in range(4):
            x = (x + np.uint64(0x9E3779B97F4A7C15)) & M
            z = x.copy()
            z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
            z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
            out[:, k] = z ^ (z >> np.uint64(31))
    return out


def default_params(kind, dtype="f32"):
    L = lib()
    if kind in (KIND_CARTPOLE, KIND_CARTPOLE_CONT):
        q = np.empty(11, dtype=np.float64)
        L.orc_cartpole_default_params(1 if dtype == "f64" else 0, _p(q))
        return q
    if kind == KIND_ACROBOT:   # AcrobotEnv.jl:19-60 (Float64): 12 fields, max_steps, book
        return np.array([1.0, 1.0, 1.0, 1.0, 0.5, 0.5, 1.0, 0.0, 4 * np.pi, 9 * np.pi, 9.8, 0.2, 200, 1], dtype=np.float64)
    T = np.float64 if dtype == "f64" else np.float32
    if kind == KIND_PENDULUM:
        return np.array([8, 2, 10, 1, 1, float(T(0.05)), 200, 3, 1], dtype=np.float64)
    f = lambda v: float(T(v))
    if kind == KIND_MOUNTAINCAR_CONT:  # MountainCarEnv.jl:73-74
        return np.array([f(-1.2), f(0.6), f(0.07), f(0.45), 0.0, f(0.0015), f(0.0025), 200], dtype=np.float64)
    return np.array([f(-1.2), f(0.6), f(0.07), f(0.5), 0.0, f(0.001), f(0.0025), 200], dtype=np.float64)


class OracleVecEnv:
    def __init__(self, kind, n, rng_states, dtype="f32", params=None):
        self.L = lib()
        self.kind, self.n, self.dtype = kind, n, dtype
        self.np_t = np.float64 if dtype == "f64" else np.float32
        self.params = default_params(kind, dtype) if params is None else np.asarray(params, np.float64)
        rng_states = np.ascontiguousarray(rng_states, dtype=np.uint64).reshape(n, 4)
        self.continuous = (kind == KIND_PENDULUM and self.params[8] != 0) or kind in (KIND_CARTPOLE_CONT, KIND_MOUNTAINCAR_CONT)
        self.h = self.L.orc_vecenv_create(kind, 1 if dtype == "f64" else 0, n, _p(self.params), _p(rng_states))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_vecenv_destroy(self.h)
            self.h = None

    def reset(self, force=True):
        self.L.orc_vecenv_reset(self.h, int(force))

    def step(self, actions, auto_reset=False):
        a = np.ascontiguousarray(actions, dtype=(self.np_t if self.continuous else np.int32))
        return self.L.orc_vecenv_step(self.h, _p(a), int(auto_reset))

    def set_max_timeout(self, max_t):
        self.L.orc_vecenv_set_max_timeout(self.h, int(max_t))

    def step_random(self, auto_reset=False):
        a = np.empty(self.n, dtype=np.int32)
        self.L.orc_vecenv_step_random(self.h, int(auto_reset), _p(a))
        return a

    def get(self, field):
        n = self.n
        shape_dt = {
            F_STATE: ((n, NS[self.kind]), self.np_t), F_OBS: ((n, NOBS[self.kind]), self.np_t),
            F_REWARD: ((n,), self.np_t), F_TERMINAL: ((n,), np.uint8), F_FLAGS: ((n,), np.uint8),
            F_T: ((n,), np.int32), F_RNG: ((n, 4), np.uint64),
            F_ACTION: ((n,), (self.np_t if self.continuous else np.float32) if (self.continuous or self.kind == KIND_PENDULUM) else np.int32),
        }[field]
        out = np.empty(*shape_dt)
        self.L.orc_vecenv_get(self.h, field, _p(out))
        return out

    def set(self, field, arr):
        arr = np.ascontiguousarray(arr)
        self.L.orc_vecenv_set(self.h, field, _p(arr))

    def bench_random(self, steps):
        return self.L.orc_vecenv_bench_random(self.h, steps)


def _mat(a, dt):
    return np.asfortranarray(np.asarray(a, dtype=dt))


def gae(rewards, values, gamma, lam, terminal=None, dims=None, dtype=np.float64):
    """generalized_advantage_estimation on Julia-shaped (column-major) inputs."""
    r = np.asarray(rewards, dtype=dtype)
    v = np.asarray(values, dtype=dtype)
    if r.ndim == 1:
        r2, v2, d = r.reshape(-1, 1), v.reshape(-1, 1), 1
        t2 = None if terminal is None else np.asarray(terminal, np.uint8).reshape(-1, 1)
    else:
        assert dims in (1, 2)
        r2, v2, d = r, v, dims
        t2 = None if terminal is None else np.asarray(terminal, np.uint8)
    rf, vf = _mat(r2, dtype), _mat(v2, dtype)
    tf = None if t2 is None else _mat(t2, np.uint8)
    out = np.zeros(rf.shape, dtype=dtype, order="F")
    fn = lib().orc_gae_f64 if dtype == np.float64 else lib().orc_gae_f32
    ct = C.c_double if dtype == np.float64 else C.c_float
    fn(_p(out), _p(rf), _p(vf), _p(tf), ct(gamma), ct(lam), rf.shape[0], rf.shape[1], d)
    return out.reshape(r.shape) if r.ndim == 1 else np.array(out)


def discount_rewards(rewards, gamma, terminal=None, init=None, dims=None, dtype=np.float64, reduced=False):
    r = np.asarray(rewards, dtype=dtype)
    if r.ndim == 1:
        r2, d = r.reshape(-1, 1), 1
        t2 = None if terminal is None else np.asarray(terminal, np.uint8).reshape(-1, 1)
        i2 = None if init is None else np.array([init], dtype=dtype)
    else:
        assert dims in (1, 2)
        r2, d = r, dims
        t2 = None if terminal is None else np.asarray(terminal, np.uint8)
        i2 = None if init is None else np.ascontiguousarray(init, dtype=dtype)
    rf = _mat(r2, dtype)
    tf = None if t2 is None else _mat(t2, np.uint8)
    ct = C.c_double if dtype == np.float64 else C.c_float
    sfx = "f64" if dtype == np.float64 else "f32"
    if reduced:
        ns = rf.shape[1] if d == 1 else rf.shape[0]
        out = np.zeros(ns, dtype=dtype)
        getattr(lib(), "orc_discount_reduced_" + sfx)(_p(out), _p(rf), _p(tf), _p(i2), ct(gamma), rf.shape[0], rf.shape[1], d)
        return out[0] if r.ndim == 1 else out
    out = np.zeros(rf.shape, dtype=dtype, order="F")
    getattr(lib(), "orc_discount_" + sfx)(_p(out), _p(rf), _p(tf), _p(i2), ct(gamma), rf.shape[0], rf.shape[1], d)
    return out.reshape(r.shape) if r.ndim == 1 else np.array(out)


# ---- learner side ---------------------------------------------------------------------------
ACT_RELU, ACT_TANH = 0, 1
HYPER_KEYS = ("gamma", "lambda_", "clip_range", "max_grad_norm", "w_actor", "w_critic", "w_entropy", "lr", "beta1", "beta2", "eps",
              "min_sigma", "max_sigma", "normalize_adv")
HYPER_DEFAULT = dict(gamma=0.99, lambda_=0.95, clip_range=0.1, max_grad_norm=0.5, w_actor=1.0, w_critic=0.5, w_entropy=0.001,
                     lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, min_sigma=0.0, max_sigma=float("inf"), normalize_adv=1)


def hyper_array(**kw):
    h = dict(HYPER_DEFAULT); h.update(kw)
    return np.array([h[k] for k in HYPER_KEYS], dtype=np.float32)


def ac_desc(n_in, hidden, n_out, act=ACT_RELU, gaussian=False):
    return np.array([n_in, hidden, act, n_out, int(gaussian)], dtype=np.int32)


def ac_nparams(desc):
    return int(lib().orc_ac_nparams(_p(desc)))


def q_nparams(desc):
    return int(lib().orc_q_nparams(_p(desc)))


def glorot_params(desc, seed, q_net=False):
    """Flux Dense default init: glorot_uniform weights U(+-sqrt(6/(in+out))), zero bias, in
    Flux.destructure order (weights (out,in) column-major).  numpy Generator stream — the
    parameters are passed in by the host, so only the shapes/order matter for parity."""
    n_in, H, _, n_out, gauss = [int(x) for x in desc]
    rng = np.random.default_rng(seed)
    def dense(o, i):
        lim = np.sqrt(6.0 / (i + o))
        return [rng.uniform(-lim, lim, (o, i)).astype(np.float32).ravel(order="F"), np.zeros(o, np.float32)]
    def mlp(heads):
        parts = dense(H, n_in) + dense(H, H)
        for d in heads:
            parts += dense(d, H)
        return parts
    if q_net:
        return np.concatenate(mlp([n_out]))
    return np.concatenate(mlp([n_out, n_out] if gauss else [n_out]) + mlp([1]))


def act_discrete(desc, params, obs, rng_states):
    n = obs.shape[1]
    obs = np.asfortranarray(obs, dtype=np.float32)
    rng_states = np.ascontiguousarray(rng_states, dtype=np.uint64).copy()
    na = int(desc[3])
    action = np.empty(n, np.int32); logp = np.empty(n, np.float32); value = np.empty(n, np.float32)
    logits = np.empty((na, n), np.float32, order="F"); margin = np.empty(n, np.float64)
    lib().orc_act_discrete(_p(desc), _p(params), _p(obs), _p(rng_states), n, _p(action), _p(logp), _p(value), _p(logits), _p(margin))
    return dict(action=action, logp=logp, value=value, logits=logits, margin=margin, rng=rng_states)


def act_gaussian(desc, hyper, params, obs, rng_states):
    n = obs.shape[1]
    obs = np.asfortranarray(obs, dtype=np.float32)
    rng_states = np.ascontiguousarray(rng_states, dtype=np.uint64).copy()
    action = np.empty(n, np.float32); logp = np.empty(n, np.float32); value = np.empty(n, np.float32)
    ms = np.empty((n, 2), np.float32)
    lib().orc_act_gaussian(_p(desc), _p(hyper), _p(params), _p(obs), _p(rng_states), n, _p(action), _p(logp), _p(value), _p(ms))
    return dict(action=action, logp=logp, value=value, mu=ms[:, 0].copy(), sigma=ms[:, 1].copy(), rng=rng_states)


def critic_values(desc, params, obs):
    obs = np.asfortranarray(obs, dtype=np.float32)
    out = np.empty(obs.shape[1], np.float32)
    lib().orc_critic_values(_p(desc), _p(params), _p(obs), obs.shape[1], _p(out))
    return out


def q_values(desc, params, obs):
    obs = np.asfortranarray(obs, dtype=np.float32)
    out = np.empty((int(desc[3]), obs.shape[1]), np.float32, order="F")
    lib().orc_q_values(_p(desc), _p(params), _p(obs), obs.shape[1], _p(out))
    return out


def explorer6(eps_stable, eps_init=1.0, warmup_steps=0, decay_steps=0, kind="linear", is_break_tie=False):
    """EpsilonGreedyExplorer(; kwargs...) as the 6 doubles the oracle takes."""
    return np.array([eps_stable, eps_init, warmup_steps, decay_steps, {"linear": 0, "exp": 1}[kind], int(is_break_tie)], np.float64)


def get_eps(ex6, step):
    return float(lib().orc_get_eps(_p(ex6), int(step)))


def egreedy_prob(ex6, step, values):
    v = np.ascontiguousarray(values, np.float64)
    out = np.empty(v.size, np.float64)
    lib().orc_egreedy_prob(_p(ex6), int(step), _p(v), v.size, _p(out))
    return out


def egreedy_plan(ex6, step0, qv, rng):
    """BatchExplorer over the columns of qv (na, N); rng (N, 4) uint64 advanced in place; 1-based actions."""
    qv = np.asfortranarray(qv, np.float32)
    na, n = qv.shape
    out = np.empty(n, np.int32)
    lib().orc_egreedy_plan(_p(ex6), int(step0), _p(qv), na, n, _p(rng), _p(out))
    return out


def ac_loss_grad(algo, desc, hyper, params, states, actions, logp_old, adv, ret, idx=None, adv_mean=0.0, adv_inv_std=1.0):
    states = np.asfortranarray(states, dtype=np.float32)
    B = len(idx) if idx is not None else states.shape[1]
    idx = None if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
    grad = np.zeros(ac_nparams(desc), np.float64); losses = np.zeros(4, np.float64)
    logp_old = None if logp_old is None else np.ascontiguousarray(logp_old, np.float32)
    lib().orc_ac_loss_grad(algo, _p(desc), _p(hyper), _p(params), _p(states), _p(np.ascontiguousarray(actions)), _p(logp_old),
                           _p(np.ascontiguousarray(adv, np.float32)), _p(np.ascontiguousarray(ret, np.float32)), _p(idx), B,
                           C.c_float(adv_mean), C.c_float(adv_inv_std), _p(grad), _p(losses))
    return grad, dict(actor_loss=losses[0], critic_loss=losses[1], entropy=losses[2], loss=losses[3])


def clip_by_global_norm(g, clip):
    g = np.ascontiguousarray(g, np.float32).copy()
    gn = lib().orc_clip_by_global_norm(_p(g), g.size, C.c_float(clip))
    return g, gn


def adam_step(p, g, m, v, beta_t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    lib().orc_adam_step(_p(p), _p(np.ascontiguousarray(g, np.float32)), _p(m), _p(v), _p(beta_t), p.size, C.c_float(lr), C.c_float(b1),
                        C.c_float(b2), C.c_float(eps))


def dqn_loss_grad(desc, p, p_target, s, a, r, t, s2, w=None, gamma=0.99, huber=True, double_dqn=False):
    s = np.asfortranarray(s, np.float32); s2 = np.asfortranarray(s2, np.float32)
    B = s.shape[1]
    grad = np.zeros(q_nparams(desc), np.float64); loss = np.zeros(1, np.float64); td = np.empty(B, np.float32)
    w = None if w is None else np.ascontiguousarray(w, np.float32)
    lib().orc_dqn_loss_grad(_p(desc), _p(p), _p(p_target), _p(s), _p(np.ascontiguousarray(a, np.int32)), _p(np.ascontiguousarray(r, np.float32)),
                            _p(np.ascontiguousarray(t, np.uint8)), _p(s2), _p(w), B, C.c_float(gamma), int(huber), int(double_dqn),
                            _p(grad), _p(loss), _p(td))
    return grad, float(loss[0]), td


def adv_norm(adv):
    adv = np.ascontiguousarray(adv, np.float32)
    out = np.empty(2, np.float32)
    lib().orc_adv_norm(_p(adv), adv.size, _p(out))
    return float(out[0]), float(out[1])


def perm_index(q, n, key):
    return int(lib().orc_perm_index(q, n, key))


class OracleTraj:
    def __init__(self, ns, lanes, cap, prioritized=False, default_priority=1.0):
        self.L = lib()
        self.ns, self.lanes, self.cap = ns, lanes, cap
        self.h = self.L.orc_traj_create(ns, lanes, cap, int(prioritized), C.c_float(default_priority))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_traj_destroy(self.h); self.h = None

    def __len__(self):
        return int(self.L.orc_traj_length(self.h))

    def push_state(self, obs):
        self.L.orc_traj_push_state(self.h, _p(np.asfortranarray(obs, np.float32)))

    def push_episode_start(self, obs, pending_only=False):
        self.L.orc_traj_push_episode_start(self.h, _p(np.asfortranarray(obs, np.float32)), int(pending_only))

    def lane_lengths(self):
        out = np.empty(self.lanes, np.int64)
        self.L.orc_traj_lane_lengths(self.h, _p(out))
        return out

    def n_sampleable(self):
        return int(self.L.orc_traj_n_sampleable(self.h))

    def push(self, a, r, t, next_obs):
        self.L.orc_traj_push(self.h, _p(np.ascontiguousarray(a, np.int32)), _p(np.ascontiguousarray(r, np.float32)),
                             _p(np.ascontiguousarray(t, np.uint8)), _p(np.asfortranarray(next_obs, np.float32)))

    def sample(self, slots, B, prioritized=False, beta=0.4):
        slots = np.ascontiguousarray(slots, np.uint64)
        s = np.empty((self.ns, B), np.float32, order="F"); s2 = np.empty((self.ns, B), np.float32, order="F")
        a = np.empty(B, np.int32); r = np.empty(B, np.float32); t = np.empty(B, np.uint8)
        key = np.empty(B, np.int64); prio = np.empty(B, np.float32); w = np.empty(B, np.float32)
        self.L.orc_traj_sample(self.h, int(prioritized), _p(slots), B, C.c_float(beta), _p(s), _p(a), _p(r), _p(t), _p(s2), _p(key), _p(prio), _p(w))
        return dict(state=s, action=a, reward=r, terminal=t, next_state=s2, key=key, priority=prio, weight=w)

    def update_priority(self, key, prio):
        key = np.ascontiguousarray(key, np.int64); prio = np.ascontiguousarray(prio, np.float32)
        self.L.orc_traj_update_priority(self.h, _p(key), _p(prio), key.size)

    def total_priority(self):
        return float(self.L.orc_traj_total_priority(self.h))


def ppo_iteration(env, desc, hyper, params, m, v, beta_t, policy_rng, T=32, E=4, M=4, update_counter=0, want_stats=False):
    """One full CPU PPO iteration on an OracleVecEnv (in-place on params / optimiser state / rng). Returns (seconds, stats)."""
    stats = np.zeros((E * M, 5), np.float32) if want_stats else None
    sec = lib().orc_ppo_iteration(env.h, _p(desc), _p(hyper), _p(params), _p(m), _p(v), _p(beta_t), _p(policy_rng), T, E, M, update_counter, _p(stats))
    return sec, stats
