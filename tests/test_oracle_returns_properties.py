"""Size-independent properties of the returns / GAE restatement (hypothesis).  These are the identities the full-size GPU test
(tests/test_returns_gpu.py::test_baseline_size_properties) relies on; here they are checked on the oracle itself, beyond the reference's
29 golden vectors, so the checker is trustworthy on shapes the vectors do not cover."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle_lib as O

SET = dict(max_examples=60, deadline=None)
shapes = st.tuples(st.integers(1, 9), st.integers(1, 17))


def _inputs(seed, R, C, dims):
    rng = np.random.default_rng(seed)
    r = rng.standard_normal((R, C))
    term = rng.random((R, C)) < 0.15
    vshape = (R + 1, C) if dims == 1 else (R, C + 1)
    return r, rng.standard_normal(vshape), term, rng


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), shape=shapes, dims=st.sampled_from([1, 2]), gamma=st.floats(0.0, 1.0))
def test_lambda_one_gae_is_discounted_return_minus_value(seed, shape, dims, gamma):
    """A(lambda = 1) = G - V with G bootstrapped from the last value (telescoping sum of the deltas)."""
    r, v, term, _ = _inputs(seed, *shape, dims)
    init = v[-1, :] if dims == 1 else v[:, -1]
    v_now = v[:-1, :] if dims == 1 else v[:, :-1]
    adv = O.gae(r, v, gamma, 1.0, terminal=term, dims=dims)
    gains = O.discount_rewards(r, gamma, terminal=term, init=init, dims=dims)
    np.testing.assert_allclose(adv, gains - v_now, rtol=1e-9, atol=1e-9)


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), shape=shapes, dims=st.sampled_from([1, 2]), gamma=st.floats(0.0, 1.0), lam=st.floats(0.0, 1.0))
def test_linearity_and_transposition(seed, shape, dims, gamma, lam):
    r, v, term, rng = _inputs(seed, *shape, dims)
    r2, v2 = rng.standard_normal(r.shape), rng.standard_normal(v.shape)
    a, b = 1.7, -0.6
    lhs = O.gae(a * r + b * r2, a * v + b * v2, gamma, lam, terminal=term, dims=dims)
    rhs = a * O.gae(r, v, gamma, lam, terminal=term, dims=dims) + b * O.gae(r2, v2, gamma, lam, terminal=term, dims=dims)
    np.testing.assert_allclose(lhs, rhs, rtol=1e-9, atol=1e-9)
    other = 3 - dims                                          # the same series laid out along the other dimension
    np.testing.assert_array_equal(O.gae(r, v, gamma, lam, terminal=term, dims=dims),
                                  O.gae(r.T, v.T, gamma, lam, terminal=term.T, dims=other).T)
    np.testing.assert_array_equal(O.discount_rewards(r, gamma, terminal=term, dims=dims),
                                  O.discount_rewards(r.T, gamma, terminal=term.T, dims=other).T)


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), shape=shapes, dims=st.sampled_from([1, 2]), gamma=st.floats(0.0, 1.0), lam=st.floats(0.0, 1.0))
def test_terminal_cuts_and_reduced_is_first_element(seed, shape, dims, gamma, lam):
    r, v, _, rng = _inputs(seed, *shape, dims)
    all_term = np.ones(r.shape, bool)
    v_now = v[:-1, :] if dims == 1 else v[:, :-1]
    np.testing.assert_array_equal(O.discount_rewards(r, gamma, terminal=all_term, dims=dims), r)           # every step ends an episode
    np.testing.assert_array_equal(O.gae(r, v, gamma, lam, terminal=all_term, dims=dims), r - v_now)
    term = rng.random(r.shape) < 0.2
    n_series = r.shape[1] if dims == 1 else r.shape[0]
    init = rng.standard_normal(n_series)
    full = O.discount_rewards(r, gamma, terminal=term, init=init, dims=dims)
    red = O.discount_rewards(r, gamma, terminal=term, init=init, dims=dims, reduced=True)
    np.testing.assert_array_equal(red, full[0, :] if dims == 1 else full[:, 0])
    # gamma = 0: the gain is the reward, the advantage the one-step TD error without bootstrap
    np.testing.assert_array_equal(O.discount_rewards(r, 0.0, terminal=term, dims=dims), r)
    np.testing.assert_array_equal(O.gae(r, v, 0.0, 0.0, terminal=term, dims=dims), r - v_now)


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), shape=shapes, dims=st.sampled_from([1, 2]))
def test_float32_path_is_the_same_recurrence_in_single_precision(seed, shape, dims):
    """The f32 oracle performs the reference's operations in Float32: a serial numpy restatement in float32 agrees bit for bit."""
    r, v, term, _ = _inputs(seed, *shape, dims)
    r, v = r.astype(np.float32), v.astype(np.float32)
    g, l = np.float32(0.99), np.float32(0.95)
    got = O.gae(r, v, g, l, terminal=term, dims=dims, dtype=np.float32)
    rr, vv, tt = (r, v, term) if dims == 1 else (r.T, v.T, term.T)      # time along axis 0
    exp = np.zeros_like(rr)
    for s in range(rr.shape[1]):
        gae = np.float32(0)
        for i in range(rr.shape[0] - 1, -1, -1):
            c = np.float32(0.0 if tt[i, s] else 1.0)
            delta = np.float32(np.float32(rr[i, s] + np.float32(np.float32(g * vv[i + 1, s]) * c)) - vv[i, s])
            gae = np.float32(delta + np.float32(np.float32(np.float32(g * l) * c) * gae))
            exp[i, s] = gae
    np.testing.assert_array_equal(got, exp if dims == 1 else exp.T)
