"""Pin the CPU oracle's GAE / discounted-return restatement to the reference's own golden vectors:
tests/golden/returns_reference_vectors.json, transcribed (tests/golden/make_returns_vectors.py) from the assertions of
/root/reference/src/ReinforcementLearningCore/test/utils/base.jl:22-152; `≈` there is isapprox with rtol = sqrt(eps)."""
import json
import os

import numpy as np
import pytest

RT = 1.5e-8
_G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "returns_reference_vectors.json")))


def _arr(x):
    if isinstance(x, dict):   # reshape(1:n, r, c): Julia column-major
        return np.arange(1, x["reshape_1_to"] + 1, dtype=np.float64).reshape(*x["shape"], order="F")
    return x


def _kw(c):
    kw = {}
    for k in ("init", "dims"):
        if k in c:
            kw[k] = c[k]
    if "terminal" in c:
        t = c["terminal"]
        kw["terminal"] = np.array(t, dtype=np.uint8) if isinstance(t[0], list) else t
    return kw


def _exp(e):
    return np.array(e, dtype=np.float64) if isinstance(e, list) and e and isinstance(e[0], list) else e


DISCOUNT_CASES = [(_arr(c["rewards"]), c["gamma"], _kw(c), _exp(c["expected"])) for c in _G["discount_rewards"]]
REDUCED_CASES = [(_arr(c["rewards"]), c["gamma"], _kw(c), _exp(c["expected"])) for c in _G["discount_rewards_reduced"]]
GAE_CASES = [(_arr(c["rewards"]), _arr(c["values"]), c["gamma"], c["lambda"], _kw(c), _exp(c["expected"]))
             for c in _G["generalized_advantage_estimation"]]
ENV_KAT = _G["env_known_answers"]


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
