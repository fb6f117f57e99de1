"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liboracle.so")

F_STATE, F_OBS, F_REWARD, F_TERMINAL, F_T, F_RNG, F_FLAGS = 0, 1, 2, 3, 4, 5, 6
KIND_CARTPOLE, KIND_PENDULUM, KIND_MOUNTAINCAR = 0, 1, 2
NS = {0: 4, 1: 2, 2: 2}
NOBS = {0: 4, 1: 3, 2: 2}


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
        "orc_cartpole_default_params": (None, [i32, vp]),
        "orc_vecenv_create": (vp, [i32, i32, i64, vp, vp]), "orc_vecenv_destroy": (None, [vp]),
        "orc_vecenv_reset": (None, [vp, i32]), "orc_vecenv_step": (i32, [vp, vp, i32]),
        "orc_vecenv_step_random": (i32, [vp, i32, vp]), "orc_vecenv_get": (None, [vp, i32, vp]),
        "orc_vecenv_set": (None, [vp, i32, vp]), "orc_vecenv_bench_random": (f64, [vp, i32]),
        "orc_gae_f32": (None, [vp, vp, vp, vp, f32, f32, i64, i64, i32]),
        "orc_gae_f64": (None, [vp, vp, vp, vp, f64, f64, i64, i64, i32]),
        "orc_discount_f32": (None, [vp, vp, vp, vp, f32, i64, i64, i32]),
        "orc_discount_f64": (None, [vp, vp, vp, vp, f64, i64, i64, i32]),
        "orc_discount_reduced_f32": (None, [vp, vp, vp, vp, f32, i64, i64, i32]),
        "orc_discount_reduced_f64": (None, [vp, vp, vp, vp, f64, i64, i64, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _L = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


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
    if kind == KIND_CARTPOLE:
        q = np.empty(11, dtype=np.float64)
        L.orc_cartpole_default_params(1 if dtype == "f64" else 0, _p(q))
        return q
    if kind == KIND_PENDULUM:
        return np.array([8, 2, 10, 1, 1, float(np.float32(0.05)), 200, 3, 1], dtype=np.float64)
    f = lambda v: float(np.float32(v))
    return np.array([f(-1.2), f(0.6), f(0.07), f(0.5), 0.0, f(0.001), f(0.0025), 200], dtype=np.float64)


class OracleVecEnv:
    def __init__(self, kind, n, rng_states, dtype="f32", params=None):
        self.L = lib()
        self.kind, self.n, self.dtype = kind, n, dtype
        self.np_t = np.float64 if dtype == "f64" else np.float32
        self.params = default_params(kind, dtype) if params is None else np.asarray(params, np.float64)
        rng_states = np.ascontiguousarray(rng_states, dtype=np.uint64).reshape(n, 4)
        self.continuous = kind == KIND_PENDULUM and self.params[8] != 0
        self.h = self.L.orc_vecenv_create(kind, 1 if dtype == "f64" else 0, n, _p(self.params), _p(rng_states))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_vecenv_destroy(self.h)
            self.h = None

    def reset(self, force=True):
        self.L.orc_vecenv_reset(self.h, int(force))

    def step(self, actions, auto_reset=False):
        a = np.ascontiguousarray(actions, dtype=np.float32 if self.continuous else np.int32)
        return self.L.orc_vecenv_step(self.h, _p(a), int(auto_reset))

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
