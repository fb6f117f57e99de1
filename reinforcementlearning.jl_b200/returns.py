"""generalized_advantage_estimation / discount_rewards / discount_rewards_reduced with the
reference's signatures (RLCore/src/utils/basic.jl:138-417), computed by the CUDA scan kernels.

Inputs are host arrays in Julia shape: vectors, or matrices with ``dims`` = 1 (time along
dim 1) or 2 (time along dim 2).  Output dtype follows the reference's
``promote_type(eltype(rewards), typeof(gamma))`` for Float32/Float64 (integers promote to the
gamma type)."""
import ctypes as C

import numpy as np

from . import _lib as L


def _dtype(rewards, gamma):
    r = np.asarray(rewards)
    g = np.float32 if isinstance(gamma, np.float32) else np.float64
    if r.dtype == np.float64 or g is np.float64:
        return np.float64
    return np.float32


def _prep(rewards, terminal, dims, dt):
    r = np.asarray(rewards, dtype=dt)
    if r.ndim == 1:
        r2 = r.reshape(-1, 1)
        t2 = None if terminal is None else np.asarray(terminal, dtype=np.uint8).reshape(-1, 1)
        d = 1
    elif r.ndim == 2:
        if dims not in (1, 2):
            raise TypeError("MethodError: for 2d rewards, the keyword argument `dims` must be either 1 or 2")
        r2, d = r, dims
        t2 = None if terminal is None else np.asarray(terminal, dtype=np.uint8)
    else:
        raise TypeError("rewards must be a vector or a matrix")
    if t2 is not None and t2.shape != r2.shape:
        raise ValueError("terminal must have the same size as rewards")
    return r, np.asfortranarray(r2), (None if t2 is None else np.asfortranarray(t2)), d


def generalized_advantage_estimation(ctx, rewards, values, gamma, lam, dims=None, terminal=None):
    if type(gamma) is not type(lam):
        raise TypeError("MethodError: gamma and lambda must have the same type")
    dt = _dtype(rewards, gamma)
    r, rf, tf, d = _prep(rewards, terminal, dims, dt)
    v = np.asarray(values, dtype=dt)
    vf = np.asfortranarray(v.reshape(-1, 1) if v.ndim == 1 else v)
    R, Cc = rf.shape
    want = (R + 1, Cc) if d == 1 else (R, Cc + 1)
    if vf.shape != want:
        raise ValueError(f"values must have shape {want}, got {vf.shape}")
    out = np.empty(rf.shape, dtype=dt, order="F")
    if rf.size == 0:   # an empty series: the reference's loop body never runs (basic.jl:408-417); nothing to launch
        return out.reshape(r.shape) if r.ndim == 1 else out
    if dt is np.float64:
        L.check(ctx.lib.b200rl_gae_f64(ctx.h, L.ptr(out), L.ptr(rf), L.ptr(vf), L.ptr(tf), C.c_double(gamma), C.c_double(lam), R, Cc, d, 0))
    else:
        L.check(ctx.lib.b200rl_gae_f32(ctx.h, L.ptr(out), L.ptr(rf), L.ptr(vf), L.ptr(tf), C.c_float(gamma), C.c_float(lam), R, Cc, d, 0))
    return out.reshape(r.shape) if r.ndim == 1 else out


def _discount(ctx, rewards, gamma, dims, terminal, init, reduced):
    dt = _dtype(rewards, gamma)
    r, rf, tf, d = _prep(rewards, terminal, dims, dt)
    R, Cc = rf.shape
    n_series = Cc if d == 1 else R
    if init is None:
        i2 = None
    elif r.ndim == 1:
        i2 = np.array([init], dtype=dt)
    else:
        i2 = np.ascontiguousarray(init, dtype=dt)
        if i2.shape != (n_series,):
            raise ValueError("init must have one entry per series")
    if rf.size == 0:   # empty input -> empty output of the promoted eltype (the reference's loops never run)
        if reduced:        # gain = init (zero when absent): basic.jl:253-263
            if r.ndim == 1:
                return dt(0) if init is None else dt(init)
            return np.zeros(n_series, dtype=dt) if i2 is None else i2.copy()
        return np.empty(r.shape, dtype=dt, order="F")
    ct = C.c_double if dt is np.float64 else C.c_float
    sfx = "f64" if dt is np.float64 else "f32"
    if reduced:
        out = np.empty(n_series, dtype=dt)
        L.check(getattr(ctx.lib, "b200rl_discount_rewards_reduced_" + sfx)(ctx.h, L.ptr(out), L.ptr(rf), L.ptr(tf), L.ptr(i2), ct(gamma), R, Cc, d, 0))
        return out[0] if r.ndim == 1 else out
    out = np.empty(rf.shape, dtype=dt, order="F")
    L.check(getattr(ctx.lib, "b200rl_discount_rewards_" + sfx)(ctx.h, L.ptr(out), L.ptr(rf), L.ptr(tf), L.ptr(i2), ct(gamma), R, Cc, d, 0))
    return out.reshape(r.shape) if r.ndim == 1 else out


def discount_rewards(ctx, rewards, gamma, dims=None, terminal=None, init=None):
    return _discount(ctx, rewards, gamma, dims, terminal, init, False)


def discount_rewards_reduced(ctx, rewards, gamma, dims=None, terminal=None, init=None):
    return _discount(ctx, rewards, gamma, dims, terminal, init, True)
