"""Pin the CPU oracle's GAE / discounted-return restatement to the reference's own golden
vectors: /root/reference/src/ReinforcementLearningCore/test/utils/base.jl:22-152 (values
copied from the assertions there; `≈` is isapprox with rtol = sqrt(eps))."""
import numpy as np
import pytest

RT = 1.5e-8


def F(x):  # Julia column-major matrix literal written row by row
    return np.array(x, dtype=np.float64)


R9 = np.arange(1, 10, dtype=np.float64).reshape(3, 3, order="F")     # reshape(1:9, 3, 3)
V43 = np.arange(1, 13, dtype=np.float64).reshape(4, 3, order="F")    # reshape(1:12, 4, 3)
V34 = np.arange(1, 13, dtype=np.float64).reshape(3, 4, order="F")    # reshape(1:12, 3, 4)
TERM = np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]], dtype=np.uint8)
INIT = [-2.0, 0.0, 2.0]

DISCOUNT_CASES = [  # (rewards, gamma, kwargs, expected)   base.jl:22-62
    ([1.0], 0.5, {}, [1.0]),
    ([1.0], 0.5, dict(init=2.0), [2.0]),
    ([1, 2, 3], 0.5, {}, [2.75, 3.5, 3.0]),
    ([1, 2, 3], 0.5, dict(init=4.0), [3.25, 4.5, 5.0]),
    ([1, 2, 3], 0.5, dict(terminal=[0, 1, 0], init=2.0), [2.0, 2.0, 4.0]),
    ([1, 2, 3], 0.5, dict(terminal=[1, 0, 1], init=2.0), [1.0, 3.5, 3.0]),
    (R9, 0.5, dict(dims=1), F([[2.75, 8.0, 13.25], [3.5, 8.0, 12.5], [3.0, 6.0, 9.0]])),
    (R9, 0.5, dict(dims=2), F([[4.75, 7.5, 7.0], [6.5, 9.0, 8.0], [8.25, 10.5, 9.0]])),
    (R9, 0.5, dict(init=INIT, dims=1), F([[2.5, 8.0, 13.5], [3.0, 8.0, 13.0], [2.0, 6.0, 10.0]])),
    (R9, 0.5, dict(init=INIT, dims=2), F([[4.5, 7.0, 6.0], [6.5, 9.0, 8.0], [8.5, 11.0, 10]])),
    (R9, 0.5, dict(dims=1, terminal=TERM), F([[2.0, 4.0, 11.0], [2.0, 8.0, 8.0], [3.0, 6.0, 9.0]])),
    (R9, 0.5, dict(dims=1, terminal=TERM, init=INIT), F([[2.0, 4.0, 11.0], [2.0, 8.0, 8.0], [2.0, 6.0, 10.0]])),
    (R9, 0.5, dict(dims=2, terminal=TERM, init=INIT), F([[3.0, 4.0, 6.0], [2.0, 9.0, 8.0], [6.0, 6.0, 10.0]])),
]
REDUCED_CASES = [  # base.jl:64-102
    ([1.0], 0.5, {}, 1.0),
    ([1, 2, 3], 0.5, {}, 2.75),
    ([1, 2, 3], 0.5, dict(init=4.0), 3.25),
    ([1, 2, 3], 0.5, dict(terminal=[0, 1, 0]), 2.0),
    ([1, 2, 3], 0.5, dict(terminal=[0, 1, 0], init=4.0), 2.0),
    (R9, 0.5, dict(dims=1), [2.75, 8.0, 13.25]),
    (R9, 0.5, dict(dims=2), [4.75, 6.5, 8.25]),
    (R9, 0.5, dict(dims=1, terminal=TERM, init=INIT), [2.0, 4.0, 11.0]),
    (R9, 0.5, dict(dims=2, terminal=TERM, init=INIT), [3.0, 2.0, 6.0]),
]
GAE_CASES = [  # (rewards, values, gamma, lambda, kwargs, expected)   base.jl:104-152
    ([1.0], [2.0, 3.0], 0.5, 0.3, {}, [0.5]),
    ([1.0, 1.0], [1, 2, 3], 0.5, 0.3, {}, [1.075, 0.5]),
    ([1, 2, 3], [1, 2, 3, 4], 0.5, 0.3, {}, [1.27, 1.8, 2]),
    ([1, 2, 3], [1, 2, 3, 4], 0.5, 0.3, dict(terminal=[1, 0, 1]), [0.0, 1.5, 0.0]),
    (R9, V43, 0.5, 0.3, dict(dims=1), F([[1.27, 2.4425, 3.615], [1.8, 2.95, 4.1], [2.0, 3.0, 4.0]])),
    (R9, V34, 0.5, 0.3, dict(dims=2), F([[2.6375, 4.25, 5.0], [3.22375, 4.825, 5.5], [3.81, 5.4, 6.0]])),
    (R9, V43, 0.5, 0.3, dict(dims=1, terminal=TERM), F([[1.0, -1.0, 2.7], [0.0, 2.35, -2.0], [2.0, -1.0, 4.0]])),
]


@pytest.mark.parametrize("r,g,kw,exp", DISCOUNT_CASES)
def test_discount_rewards_golden(oracle, r, g, kw, exp):
    np.testing.assert_allclose(oracle.discount_rewards(r, g, **kw), exp, rtol=RT)


@pytest.mark.parametrize("r,g,kw,exp", REDUCED_CASES)
def test_discount_rewards_reduced_golden(oracle, r, g, kw, exp):
    np.testing.assert_allclose(oracle.discount_rewards(r, g, reduced=True, **kw), exp, rtol=RT)


@pytest.mark.parametrize("r,v,g,l,kw,exp", GAE_CASES)
def test_gae_golden(oracle, r, v, g, l, kw, exp):
    np.testing.assert_allclose(oracle.gae(r, v, g, l, **kw), exp, rtol=RT)


def test_gae_float32_matches_float64_formula(oracle):
    rng = np.random.default_rng(3)
    r = rng.standard_normal((17, 9)).astype(np.float32)
    v = rng.standard_normal((17, 10)).astype(np.float32)
    t = (rng.random((17, 9)) < 0.2).astype(np.uint8)
    a32 = oracle.gae(r, v, 0.99, 0.95, terminal=t, dims=2, dtype=np.float32)
    a64 = oracle.gae(r, v, float(np.float32(0.99)), float(np.float32(0.95)), terminal=t, dims=2, dtype=np.float64)
    np.testing.assert_allclose(a32, a64, rtol=2e-5, atol=2e-6)
