"""GPU parity of K5 (GAE / discounted returns scans) through the C ABI: the reference's golden
vectors (RLCore/test/utils/base.jl:22-152) and bit-exact agreement with the CPU oracle on
random inputs in both layouts, plus size-independent properties at BASELINE sizes."""
import numpy as np
import pytest

import oracle_lib as O
from test_oracle_returns import DISCOUNT_CASES, GAE_CASES, REDUCED_CASES, RT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("r,g,kw,exp", DISCOUNT_CASES)
def test_discount_rewards_golden(pkg, ctx, r, g, kw, exp):
    np.testing.assert_allclose(pkg.discount_rewards(ctx, r, g, **kw), exp, rtol=RT)


@pytest.mark.parametrize("r,g,kw,exp", REDUCED_CASES)
def test_discount_rewards_reduced_golden(pkg, ctx, r, g, kw, exp):
    np.testing.assert_allclose(pkg.discount_rewards_reduced(ctx, r, g, **kw), exp, rtol=RT)


@pytest.mark.parametrize("r,v,g,l,kw,exp", GAE_CASES)
def test_gae_golden(pkg, ctx, r, v, g, l, kw, exp):
    np.testing.assert_allclose(pkg.generalized_advantage_estimation(ctx, r, v, g, l, **kw), exp, rtol=RT)


def test_type_rules(pkg, ctx):
    # base.jl:41-42,123-124: Int rewards + Float32 gamma -> Float32
    assert pkg.discount_rewards(ctx, [1, 2, 3], np.float32(0.5)).dtype == np.float32
    assert pkg.generalized_advantage_estimation(ctx, [1, 2, 3], [1, 2, 3, 4], np.float32(0.5), np.float32(0.5)).dtype == np.float32
    with pytest.raises(TypeError):  # 2-D rewards need dims (MethodError in the reference)
        pkg.discount_rewards(ctx, np.ones((3, 3)), 0.5)
    with pytest.raises(TypeError):
        pkg.generalized_advantage_estimation(ctx, np.ones((3, 3)), np.ones((4, 3)), 0.5, 0.3)


def _same(a, b, exact, dt):
    """bit-exact where the serial thread-per-series kernel runs; few long series (S < 1024 and >= 64 time steps) take the warp-segmented
    scan, which re-associates the recurrence: 1e-5 relative to the scale of the series (an element next to a zero crossing of a
    discounted sum of O(10) carries the rounding of its O(10) neighbours) — in practice ~2e-6"""
    if exact:
        return np.asfortranarray(a).tobytes(order="F") == np.asfortranarray(b).tobytes(order="F")
    tol = 1e-5 if dt == np.float32 else 1e-13
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max()) <= tol * max(1.0, float(np.abs(b).max()))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape,dims", [((1000, 33), 2), ((33, 1000), 1), ((7, 5), 1), ((7, 5), 2), ((1, 300), 2), ((300, 1), 1), ((129, 65), 1),
                                        ((4, 5000), 2), ((2000, 64), 2), ((63, 900), 1)])
def test_random_vs_oracle(pkg, ctx, dt, shape, dims):
    rng = np.random.default_rng(hash((shape, dims)) & 0xFFFF)
    r = rng.standard_normal(shape).astype(dt)
    vshape = (shape[0] + 1, shape[1]) if dims == 1 else (shape[0], shape[1] + 1)
    v = rng.standard_normal(vshape).astype(dt)
    term = (rng.random(shape) < 0.1).astype(np.uint8)
    ns, nt = (shape[1], shape[0]) if dims == 1 else (shape[0], shape[1])
    exact = not (ns < 1024 and nt >= 64)
    init = rng.standard_normal(ns).astype(dt)
    g, l = dt(0.99), dt(0.95)
    for t in (None, term):
        a = pkg.generalized_advantage_estimation(ctx, r, v, g, l, dims=dims, terminal=t)
        b = O.gae(r, v, g, l, terminal=t, dims=dims, dtype=dt)
        assert a.dtype == dt and a.shape == b.shape and _same(a, b, exact, dt)
        for ini in (None, init):
            a = pkg.discount_rewards(ctx, r, g, dims=dims, terminal=t, init=ini)
            b = O.discount_rewards(r, g, terminal=t, init=ini, dims=dims, dtype=dt)
            assert _same(a, b, exact, dt)
            a = pkg.discount_rewards_reduced(ctx, r, g, dims=dims, terminal=t, init=ini)
            b = O.discount_rewards(r, g, terminal=t, init=ini, dims=dims, dtype=dt, reduced=True)
            assert _same(a, b, exact, dt)


def test_baseline_size_properties(pkg, ctx):
    """C2 size (65536 x 32, Float32): linearity in rewards when lambda = 1 and no terminals
    (GAE(lambda=1) == discounted return - value), and terminal masking cuts the chain."""
    n, T = 65536, 32
    rng = np.random.default_rng(0)
    r = rng.random((n, T)).astype(np.float32)
    v = rng.standard_normal((n, T + 1)).astype(np.float32)
    g = np.float32(0.99)
    adv = pkg.generalized_advantage_estimation(ctx, r, v, g, np.float32(1.0), dims=2)
    ret = pkg.discount_rewards(ctx, r, g, dims=2, init=v[:, T].copy())
    np.testing.assert_allclose(adv, ret - v[:, :T], rtol=2e-4, atol=2e-4)
    term = np.zeros((n, T), np.uint8); term[:, 15] = 1
    adv_t = pkg.generalized_advantage_estimation(ctx, r, v, g, np.float32(0.95), dims=2, terminal=term)
    adv_head = pkg.generalized_advantage_estimation(ctx, r[:, :16], np.concatenate([v[:, :16], np.zeros((n, 1), np.float32)], 1), g,
                                                    np.float32(0.95), dims=2)
    assert np.array_equal(adv_t[:, :16], adv_head)
    sub = slice(0, 2048)  # oracle on a slice that finishes in well under a second
    assert np.array_equal(adv_t[sub], O.gae(r[sub], v[sub], g, np.float32(0.95), terminal=term[sub], dims=2, dtype=np.float32))
