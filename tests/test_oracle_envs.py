"""Pin the oracle's env restatement: SURVEY Appendix C known answers (derived from the
formulas at CartPoleEnv.jl:118-140), the reference's conformance properties
(RLBase/src/base.jl:86-152: same-seed copies stay identical, states stay in state_space) and
RNG-consumption rules (Appendix A.4).  The reference itself has no golden trajectories."""
import numpy as np
import pytest

import oracle_lib as O


def _fresh(kind, n, dtype="f32", seed=1, **kw):
    return O.OracleVecEnv(kind, n, O.splitmix_states_fast(n, seed), dtype=dtype, **kw)


def test_cartpole_f64_known_answers():
    e = _fresh(O.KIND_CARTPOLE, 1, "f64")
    e.set(O.F_STATE, np.zeros((1, 4)))
    e.set(O.F_T, np.zeros(1, np.int32))
    e.step([2])
    assert e.get(O.F_STATE)[0].tolist() == [0.0, 0.1951219512195122, 0.0, -0.2926829268292683]
    e.step([2]); e.step([2])
    assert e.get(O.F_STATE)[0].tolist() == [0.011707317073170733, 0.585447355516259, -0.0175609756097561, -0.8798869825388553]
    assert e.get(O.F_REWARD)[0] == 1.0 and e.get(O.F_TERMINAL)[0] == 0 and e.get(O.F_T)[0] == 3


def test_cartpole_f32_known_answers_bits():
    e = _fresh(O.KIND_CARTPOLE, 1)
    e.set(O.F_STATE, np.zeros((1, 4), np.float32))
    e.set(O.F_T, np.zeros(1, np.int32))
    e.step([2])
    assert e.get(O.F_STATE)[0].view(np.uint32).tolist() == [0x00000000, 0x3E47CE0C, 0x00000000, 0xBE95DA89]
    e.step([2]); e.step([2])
    assert e.get(O.F_STATE)[0].view(np.uint32).tolist() == [0x3C3FD00B, 0x3F15DFE0, 0xBC8FDC08, 0xBF614045]


def test_cartpole_threshold_and_time_limit():
    q = O.default_params(O.KIND_CARTPOLE, "f32")
    assert q[8] == 0.20943951606750488  # Float32(12*pi/180)
    # time limit: done at t = 201 (t > max_steps), reward of the terminating step is 0
    e = _fresh(O.KIND_CARTPOLE, 1)
    for k in range(1, 400):
        e.set(O.F_STATE, np.zeros((1, 4), np.float32))  # keep it balanced
        e.step([1 + (k % 2)])
        if e.get(O.F_TERMINAL)[0]:
            break
    assert e.get(O.F_T)[0] == 201 and e.get(O.F_REWARD)[0] == 0.0


def test_pendulum_mountaincar_time_limit_is_geq():
    for kind, act in ((O.KIND_PENDULUM, np.float32(0.0)), (O.KIND_MOUNTAINCAR, 2)):
        e = _fresh(kind, 1)
        n = 0
        while not e.get(O.F_TERMINAL)[0]:
            e.step([act]); n += 1
            assert n <= 200
        assert e.get(O.F_T)[0] == 200 or kind == O.KIND_MOUNTAINCAR


def test_reset_consumes_rng_like_reference():
    """CartPole reset! = rand(rng,T,4) [2 u64 for Float32, 4 for Float64] + rand(OneTo(2)) [1 u64]."""
    L = O.lib()
    for dtype, ndraws in (("f32", 3), ("f64", 5)):
        seeds = O.splitmix_states_fast(1, 7)
        e = O.OracleVecEnv(O.KIND_CARTPOLE, 1, seeds, dtype=dtype)  # constructor resets once
        s = seeds[0].copy()
        for _ in range(ndraws):
            L.orc_rng_next(O._p(s))
        assert np.array_equal(e.get(O.F_RNG)[0], s)
        st = e.get(O.F_STATE)[0]
        assert np.all(st >= -0.05) and np.all(st < 0.05)
    # Pendulum: 2 draws, theta in [-2pi, 0), thetadot in [-2, 0); MountainCar: 1 draw, x in [-0.6,-0.4)
    e = _fresh(O.KIND_PENDULUM, 64)
    st = e.get(O.F_STATE)
    assert np.all(st[:, 0] >= -2 * np.pi - 1e-6) and np.all(st[:, 0] <= 0) and np.all(st[:, 1] >= -2) and np.all(st[:, 1] <= 0)
    e = _fresh(O.KIND_MOUNTAINCAR, 64)
    st = e.get(O.F_STATE)
    assert np.all(st[:, 0] >= -0.6 - 1e-7) and np.all(st[:, 0] < -0.4) and np.all(st[:, 1] == 0)


def test_rand_oneto_is_top_bit_for_n2_and_uniform_for_n3():
    L = O.lib()
    s = O.splitmix_states_fast(1, 11)[0].copy()
    s2 = s.copy()
    for _ in range(200):
        u = L.orc_rng_next(O._p(s))
        a = L.orc_rng_oneto(O._p(s2), 2)
        assert a == (u >> 63) + 1
    counts = np.zeros(4, int)
    for _ in range(3000):
        counts[L.orc_rng_oneto(O._p(s), 3)] += 1
    assert counts[0] == 0 and counts[1:].min() > 850


def test_same_seed_copies_stay_identical_and_in_space():
    """RLBase.test_interfaces! determinism + state in state_space (base.jl:86-152), 1000 random steps (test_runnable!)."""
    for kind in (O.KIND_CARTPOLE, O.KIND_MOUNTAINCAR):
        a, b = _fresh(kind, 8, seed=888), _fresh(kind, 8, seed=888)
        for _ in range(1000):
            a.step_random(auto_reset=True); b.step_random(auto_reset=True)
            assert np.array_equal(a.get(O.F_STATE), b.get(O.F_STATE))
        st = a.get(O.F_STATE)
        if kind == O.KIND_CARTPOLE:
            assert np.all(np.abs(st[:, 0]) <= 4.8) and np.all(np.abs(st[:, 2]) <= 2 * 0.20943951606750488)
        else:
            assert np.all(st[:, 0] >= np.float32(-1.2)) and np.all(st[:, 0] <= np.float32(0.6)) and np.all(np.abs(st[:, 1]) <= np.float32(0.07))


def test_soft_reset_and_auto_reset_agree():
    """Fused auto-reset == step, then soft reset at the top of the next iteration (same RNG order)."""
    n = 32
    a, b = _fresh(O.KIND_CARTPOLE, n, seed=5), _fresh(O.KIND_CARTPOLE, n, seed=5)
    rng = np.random.default_rng(0)
    for _ in range(400):
        act = rng.integers(1, 3, n)
        a.step(act, auto_reset=True)
        b.reset(force=False)
        b.step(act, auto_reset=False)
        assert np.array_equal(a.get(O.F_REWARD), b.get(O.F_REWARD))
        assert np.array_equal(a.get(O.F_TERMINAL), b.get(O.F_TERMINAL))
    b.reset(force=False)
    assert np.array_equal(a.get(O.F_STATE), b.get(O.F_STATE))
    assert np.array_equal(a.get(O.F_RNG), b.get(O.F_RNG))


def test_pendulum_reward_and_obs_formulas():
    e = _fresh(O.KIND_PENDULUM, 16, seed=3)
    st0 = e.get(O.F_STATE).astype(np.float64)
    act = np.linspace(-2, 2, 16).astype(np.float32)
    e.step(act)
    th, thd = st0[:, 0], st0[:, 1]
    an = np.mod(np.float32(th.astype(np.float32) + np.float32(np.pi)).astype(np.float64), 2 * np.pi) - np.pi
    cost = an ** 2 + 0.1 * thd ** 2 + 0.001 * act.astype(np.float64) ** 2
    np.testing.assert_allclose(e.get(O.F_REWARD), -cost, rtol=1e-6, atol=1e-6)
    st = e.get(O.F_STATE).astype(np.float64)
    obs = e.get(O.F_OBS)
    np.testing.assert_allclose(obs[:, 0], np.sin(st[:, 0]), atol=1e-6)
    np.testing.assert_allclose(obs[:, 1], np.cos(st[:, 0]), atol=1e-6)
    assert np.array_equal(obs[:, 2], e.get(O.F_STATE)[:, 1])


def test_invalid_actions_are_rejected():
    e = _fresh(O.KIND_CARTPOLE, 4)
    assert e.step([1, 2, 3, 0]) == 2
    p = _fresh(O.KIND_PENDULUM, 2)
    assert p.step(np.array([2.5, np.nan], np.float32)) == 2


def test_continuous_cartpole_matches_discrete_at_unit_force_and_kat():
    """CartPoleEnv(continuous = true): _step!(env, a) with force = a * forcemag (CartPoleEnv.jl:106-110,118-120).
    a = +-1.0 is the discrete env's action 2 / 1 bit for bit; reset! consumes the same three 64-bit draws."""
    n = 64
    seeds = O.splitmix_states_fast(n, 21)
    d = O.OracleVecEnv(O.KIND_CARTPOLE, n, seeds)
    c = O.OracleVecEnv(O.KIND_CARTPOLE_CONT, n, seeds)
    assert np.array_equal(d.get(O.F_RNG), c.get(O.F_RNG)) and np.array_equal(d.get(O.F_STATE), c.get(O.F_STATE))
    rng = np.random.default_rng(0)
    for _ in range(250):
        a = rng.integers(1, 3, n).astype(np.int32)
        d.step(a, auto_reset=True)
        assert c.step(np.where(a == 2, 1.0, -1.0).astype(np.float32), auto_reset=True) == 0
        for f in (O.F_STATE, O.F_REWARD, O.F_FLAGS, O.F_T, O.F_RNG):
            assert np.array_equal(d.get(f), c.get(f))
    # known answer from the formula at the zero state, a = 0.5 (Float32 env; `4/3` sub-expressions Float64)
    e = O.OracleVecEnv(O.KIND_CARTPOLE_CONT, 1, O.splitmix_states_fast(1, 3))
    e.set(O.F_STATE, np.zeros((1, 4), np.float32)); e.set(O.F_T, np.zeros(1, np.int32))
    assert e.step(np.array([0.5], np.float32)) == 0
    f32 = np.float32
    q = O.default_params(O.KIND_CARTPOLE, "f32")
    M, m, l, pml, fmag, dt = f32(q[3]), f32(q[2]), f32(q[4]), f32(q[5]), f32(q[6]), f32(q[7])
    tmp = f32(f32(0.5) * fmag) / M
    den = float(l) * (4.0 / 3.0 - float(f32(m * f32(1)) / M))
    thacc = float(f32(f32(0) - f32(1) * tmp)) / den
    xacc = float(tmp) - (float(pml) * thacc) * 1.0 / float(M)
    want = np.array([0, f32(float(dt) * xacc), 0, f32(float(dt) * thacc)], np.float32)
    assert np.array_equal(e.get(O.F_STATE)[0].view(np.uint32), want.view(np.uint32))
    # out-of-interval / NaN actions fail the `@assert a in action_space(env)`
    assert e.step(np.array([1.0000001], np.float32)) == 1 and e.step(np.array([np.nan], np.float32)) == 1


def test_continuous_mountaincar_defaults_and_equivalence():
    """ContinuousMountainCarEnv: goal_pos 0.45, power 0.0015 (MountainCarEnv.jl:73-74); with the discrete
    params a force of a - 2 as Float32 reproduces the discrete env bit for bit."""
    q = O.default_params(O.KIND_MOUNTAINCAR_CONT)
    assert q[3] == float(np.float32(0.45)) and q[5] == float(np.float32(0.0015))
    n = 50
    seeds = O.splitmix_states_fast(n, 8)
    d = O.OracleVecEnv(O.KIND_MOUNTAINCAR, n, seeds)
    c = O.OracleVecEnv(O.KIND_MOUNTAINCAR_CONT, n, seeds, params=O.default_params(O.KIND_MOUNTAINCAR))
    rng = np.random.default_rng(1)
    for _ in range(450):
        a = rng.integers(1, 4, n).astype(np.int32)
        d.step(a, auto_reset=True); c.step((a - 2).astype(np.float32), auto_reset=True)
        for f in (O.F_STATE, O.F_REWARD, O.F_FLAGS, O.F_T, O.F_RNG):
            assert np.array_equal(d.get(f), c.get(f))
    # default continuous params: full throttle right from the valley reaches the goal before the time limit
    e = O.OracleVecEnv(O.KIND_MOUNTAINCAR_CONT, 1, O.splitmix_states_fast(1, 2))
    x0 = e.get(O.F_STATE)[0]
    assert -0.6 <= x0[0] < -0.4 and x0[1] == 0
    e.step(np.array([1.0], np.float32))
    v1 = np.float32(np.float32(0) + (np.float32(1.0) * np.float32(0.0015) + np.float32(np.cos(np.float64(np.float32(3.0) * x0[0]))) * -np.float32(0.0025)))
    assert abs(float(e.get(O.F_STATE)[0][1]) - float(v1)) <= 1e-9 and e.get(O.F_REWARD)[0] == -1.0


def test_max_timeout_wrapper_counts_like_the_reference_test():
    """RLEnvs/test/environments/wrappers/wrappers.jl:36-52: with MaxTimeoutEnv(env, n) the loop
    `while !is_terminated(env′); act!; n -= 1; @test n >= 0` ends after exactly n interactions."""
    for kind, max_t in ((O.KIND_PENDULUM, 100), (O.KIND_MOUNTAINCAR, 37)):
        e = _fresh(kind, 1)
        e.set_max_timeout(max_t)
        k = 0
        act = np.float32(0.0) if kind == O.KIND_PENDULUM else 2
        while not e.get(O.F_TERMINAL)[0]:
            e.step([act]); k += 1
            assert k <= max_t
        assert k == max_t and e.get(O.F_T)[0] == max_t
        if kind == O.KIND_MOUNTAINCAR:
            assert e.get(O.F_REWARD)[0] == -1.0    # reward(env) forwards to the wrapped env, which is not done
        e.reset(force=False)                       # reset!(env′): current_t = 1, not terminated
        assert e.get(O.F_TERMINAL)[0] == 0 and e.get(O.F_T)[0] == 0
    # wrapped CartPole: the inner termination still wins and pays 0
    e = _fresh(O.KIND_CARTPOLE, 256); e.set_max_timeout(150)
    for _ in range(150):
        e.step_random(auto_reset=False)
    assert e.get(O.F_TERMINAL).all()


# ---------------------------------------------------------------------------------------------------------------------
# AcrobotEnv{Float64} (3rd_party/AcrobotEnv.jl).  The reference holds no vectors for it and integrates with an adaptive
# OrdinaryDiffEq solver that is not in the tree; the oracle takes one classical RK4 step per act! (gym's rk4, which the
# file cites).  Pinned here: an independent NumPy restatement of gym's dsdt / rk4, the reset formula, the termination rule,
# wrap / bound and the observation.
def _gym_acrobot_step(s, a, book=True, dt=0.2):
    m1 = m2 = 1.0; l1 = 1.0; lc1 = lc2 = 0.5; I1 = I2 = 1.0; g = 9.8

    def dsdt(y):
        th1, th2, d1_, d2_ = y
        d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * np.cos(th2)) + I1 + I2
        d2 = m2 * (lc2 ** 2 + l1 * lc2 * np.cos(th2)) + I2
        phi2 = m2 * lc2 * g * np.cos(th1 + th2 - np.pi / 2.0)
        phi1 = -m2 * l1 * lc2 * d2_ ** 2 * np.sin(th2) - 2 * m2 * l1 * lc2 * d2_ * d1_ * np.sin(th2) + (m1 * lc1 + m2 * l1) * g * np.cos(th1 - np.pi / 2) + phi2
        if book:
            dd2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * d1_ ** 2 * np.sin(th2) - phi2) / (m2 * lc2 ** 2 + I2 - d2 ** 2 / d1)
            dd1 = -(d2 * dd2 + phi1) / d1
        else:
            dd2 = (a + d2 / d1 * phi1 - phi2) / (m2 * lc2 ** 2 + I2 - d2 ** 2 / d1)
            dd1 = 0.0
        return np.array([d1_, d2_, dd1, dd2])
    y0 = np.asarray(s, np.float64)
    k1 = dsdt(y0); k2 = dsdt(y0 + dt / 2 * k1); k3 = dsdt(y0 + dt / 2 * k2); k4 = dsdt(y0 + dt * k3)
    ns = y0 + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)

    def wrap(x, m, M):
        while x > M:
            x -= M - m
        while x < m:
            x += M - m
        return x
    ns[0] = wrap(ns[0], -np.pi, np.pi); ns[1] = wrap(ns[1], -np.pi, np.pi)
    ns[2] = min(max(ns[2], -4 * np.pi), 4 * np.pi); ns[3] = min(max(ns[3], -9 * np.pi), 9 * np.pi)
    return ns


@pytest.mark.parametrize("book", [1, 0])
def test_acrobot_oracle_against_an_independent_gym_restatement(book):
    n, steps = 64, 260
    seeds = O.splitmix_states_fast(n, 404)
    q = O.default_params(O.KIND_ACROBOT); q[13] = book
    e = O.OracleVecEnv(O.KIND_ACROBOT, n, seeds, dtype="f64", params=q)
    s0 = e.get(O.F_STATE)
    assert s0.shape == (n, 4) and np.all(np.abs(s0) <= 0.05) and np.all(e.get(O.F_REWARD) == -1.0)      # reset!: 0.1 rand - 0.05, reward = -1
    assert np.array_equal(e.get(O.F_ACTION), np.full(n, 2, np.int32)) and not e.get(O.F_TERMINAL).any()
    r = np.random.default_rng(5)
    state = s0.copy(); t = np.zeros(n, int)
    succeeded_seen = timeouts = 0
    for k in range(steps):
        a = r.integers(1, 4, n).astype(np.int32)
        assert e.step(a, auto_reset=False) == 0
        got, term, rew = e.get(O.F_STATE), e.get(O.F_TERMINAL), e.get(O.F_REWARD)
        obs = e.get(O.F_OBS)
        for i in range(n):
            if t[i] < 0:
                continue
            ns = _gym_acrobot_step(state[i], float(a[i] - 2), bool(book))
            t[i] += 1
            np.testing.assert_allclose(got[i], ns, rtol=1e-11, atol=1e-12)
            np.testing.assert_allclose(obs[i], [np.cos(got[i][0]), np.sin(got[i][0]), np.cos(got[i][1]), np.sin(got[i][1]), got[i][2], got[i][3]], rtol=0, atol=1e-15)
            succ = -np.cos(got[i][0]) - np.cos(got[i][1] + got[i][0]) > 1.0
            assert bool(term[i]) == (succ or t[i] > 200) and rew[i] == (0.0 if succ else -1.0)
            assert abs(got[i][0]) <= np.pi and abs(got[i][1]) <= np.pi and abs(got[i][2]) <= 4 * np.pi and abs(got[i][3]) <= 9 * np.pi
            state[i] = got[i]
            if term[i]:
                succeeded_seen += int(succ); timeouts += int(not succ)
                t[i] = -1          # finished: the oracle keeps stepping it, the check stops here
    assert timeouts > 0
    assert e.step(np.full(n, 4, np.int32)) == n          # 4 is outside Base.OneTo(3)


def test_acrobot_swings_up_with_an_energy_pumping_policy():
    """Torque in the direction of the second joint's velocity pumps energy: the swing-up termination (-cos th1 - cos(th1 + th2) > 1)
    must fire well inside 200 steps for most envs, with reward 0 on that step and -1 before."""
    n = 256
    e = O.OracleVecEnv(O.KIND_ACROBOT, n, O.splitmix_states_fast(n, 9), dtype="f64")
    done = np.zeros(n, bool); first_zero = np.full(n, -1)
    for k in range(200):
        s = e.get(O.F_STATE)
        a = np.where(s[:, 3] >= 0, 3, 1).astype(np.int32)
        e.step(a)
        rew, term = e.get(O.F_REWARD), e.get(O.F_TERMINAL).astype(bool)
        newly = term & ~done
        assert np.all(rew[newly] == 0.0) and np.all(rew[~term & ~done] == -1.0)
        first_zero[newly] = k
        done |= term
    assert done.mean() > 0.9 and np.median(first_zero[done]) < 150
