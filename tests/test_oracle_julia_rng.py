"""Known answers for the Julia stdlib `Random` pieces the oracle restates (oracle/jl_rng.hpp).  `Random` is not part of the reference
tree, so these come from the outputs Julia's own manual prints in its docstrings (stdlib/Random: `Xoshiro`, `Random.seed!`, `shuffle`,
`randperm`, `randcycle`, `randstring`; Julia 1.7–1.10, the series the reference's devcontainer pins with `julia:1.10`).  They pin the
generator core, the integer seeding, the Float64 sampler and the UInt52 masked-rejection sampler behind `shuffle!` — the same `next()`
every other sampler of the envs and policies is built on."""
import numpy as np

import oracle_lib as O


def f64(state):
    return (int(O.lib().orc_rng_next(O._p(state))) >> 11) * 2.0 ** -53


def test_xoshiro_1234_float64_vector():
    """`rng = Xoshiro(1234); x1 = rand(rng, 2)` -> [0.32597672886359486, 0.5490511363155669]   (docstring of Xoshiro / Random.seed!)"""
    s = O.julia_xoshiro(1234)
    assert [f64(s), f64(s)] == [0.32597672886359486, 0.5490511363155669]
    s2 = O.julia_xoshiro(1234)                                 # "x1 == x2": same seed, same stream
    assert f64(s2) == 0.32597672886359486


def test_shuffle_randperm_randcycle_vectors():
    """`shuffle(Xoshiro(123), Vector(1:10))` -> [5, 4, 2, 3, 6, 10, 8, 1, 9, 7]; `randperm(Xoshiro(123), 4)` -> [1, 4, 2, 3];
    `randcycle(Xoshiro(123), 6)` -> [5, 4, 2, 6, 3, 1]   (docstrings of shuffle / randperm / randcycle)"""
    assert O.jl_shuffle(O.julia_xoshiro(123), np.arange(1, 11, dtype=np.int64)).tolist() == [5, 4, 2, 3, 6, 10, 8, 1, 9, 7]
    assert O.jl_shuffle(O.julia_xoshiro(123), np.arange(1, 11, dtype=np.int32)).tolist() == [5, 4, 2, 3, 6, 10, 8, 1, 9, 7]
    assert O.jl_randperm(O.julia_xoshiro(123), 4).tolist() == [1, 4, 2, 3]
    # randcycle!: a[1] = 1; for i = 2:n  j = 1 + rand(ltm52(i - 1, mask)); a[i] = a[j]; a[j] = i
    s = O.julia_xoshiro(123)
    a, mask = [1] + [0] * 5, 3
    for i in range(2, 7):
        j = 1 + int(O.lib().orc_jl_ltm52(O._p(s), i - 1, mask))
        a[i - 1] = a[j - 1]; a[j - 1] = i
        if i == 1 + mask:
            mask = 2 * mask + 1
    assert a == [5, 4, 2, 6, 3, 1]


def test_randstring_vector_exercises_the_range_sampler():
    """`randstring(Xoshiro(3), 'a':'z', 6)` -> "iyzcsm" (docstring of randstring, recalled): six draws of rand(rng, 1:26), the
    nearly-divisionless range sampler (SamplerRangeNDL) every `rand(rng, Base.OneTo(n))` of the envs and RandomPolicy uses."""
    s = O.julia_xoshiro(3)
    assert "".join(chr(ord("a") + int(O.lib().orc_rng_oneto(O._p(s), 26)) - 1) for _ in range(6)) == "iyzcsm"


def test_ppo_minibatch_permutation_is_a_julia_shuffle():
    """`perm = shuffle!(rng, collect(1:N*T))` of the PPO update: a bijection, reproducible from the raw state, usable as perm_host."""
    n = 4096
    p = O.jl_shuffle(O.julia_xoshiro(2024), np.arange(n, dtype=np.int32))
    assert sorted(p.tolist()) == list(range(n)) and p.tolist() != list(range(n))
    assert np.array_equal(p, O.jl_shuffle(O.julia_xoshiro(2024), np.arange(n, dtype=np.int32)))
