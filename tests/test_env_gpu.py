"""GPU parity of K1/K2 (batched env step / reset / fused random policy) against the CPU
oracle through the C ABI: bit-exact state, reward, terminal flags, step counters and RNG
streams on the same seeds.  Also restates the reference's conformance tests
(RLBase/src/base.jl:86-152, 202-227) for the batched env."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

KINDS = {"CartPole": O.KIND_CARTPOLE, "Pendulum": O.KIND_PENDULUM, "MountainCar": O.KIND_MOUNTAINCAR,
         "ContinuousCartPole": O.KIND_CARTPOLE_CONT, "ContinuousMountainCar": O.KIND_MOUNTAINCAR_CONT, "Acrobot": O.KIND_ACROBOT}


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def assert_same(env, ref, what=""):
    assert np.array_equal(bits(env.internal_state().T), bits(ref.get(O.F_STATE))), f"state {what}"
    assert np.array_equal(bits(env.state().T), bits(ref.get(O.F_OBS))), f"obs {what}"
    assert np.array_equal(bits(env.reward()), bits(ref.get(O.F_REWARD))), f"reward {what}"
    assert np.array_equal(env.flags(), ref.get(O.F_FLAGS)), f"flags {what}"
    assert np.array_equal(env.t(), ref.get(O.F_T)), f"t {what}"
    assert np.array_equal(env.rng_state(), ref.get(O.F_RNG)), f"rng {what}"


@pytest.mark.parametrize("kind,T,n,steps", [
    ("CartPole", np.float32, 65536, 300), ("CartPole", np.float64, 4097, 300),
    ("MountainCar", np.float32, 10000, 450), ("CartPole", np.float32, 1, 1000), ("CartPole", np.float32, 255, 64),
    ("Acrobot", np.float64, 5000, 450),     # AcrobotEnv{Float64}: RK4 step, 6 observations, time-limit and swing-up terminations
])
def test_random_policy_autoreset_bit_exact(pkg, ctx, kind, T, n, steps):
    seeds = O.splitmix_states_fast(n, seed=0x1234 + n)
    env = pkg.B200VecEnv(ctx, kind, n, seeds, T=T, auto_reset=True)
    ref = O.OracleVecEnv(KINDS[kind], n, seeds, dtype="f64" if T is np.float64 else "f32")
    assert_same(env, ref, "after construction")
    for s in range(steps):
        env.act_random_()
        ref.step_random(auto_reset=True)
        if s % 97 == 0 or s == steps - 1:
            assert_same(env, ref, f"step {s}")
    assert np.array_equal(env.last_action(), ref.get(O.F_ACTION)), "env.action (the policy's action, or reset!'s redraw after an auto-reset)"
    st = env.episode_stats()
    assert st["episodes"] > 0 or kind == "MountainCar"
    env.close()


def test_config1_single_env_float64_1000_steps(pkg, ctx):
    """BASELINE config 1: RandomPolicy + CartPoleEnv (Float64), StopAfterNSteps(1000), one shared stream."""
    seeds = O.splitmix_states(1, seed=123)
    env = pkg.B200VecEnv(ctx, "CartPole", 1, seeds, T=np.float64, auto_reset=False)
    ref = O.OracleVecEnv(O.KIND_CARTPOLE, 1, seeds, dtype="f64")
    hook = pkg.BatchStepsPerEpisode(1)
    pkg.run(pkg.RandomPolicy(), env, pkg.StopAfterNSteps(1000), hook)
    # oracle: the same control flow written out (forced reset, then soft reset / plan / act)
    ref.reset(force=True)
    lens, cur = [], 0
    for _ in range(1000):
        ref.reset(force=False)
        ref.step_random(auto_reset=False)
        cur += 1
        if ref.get(O.F_TERMINAL)[0]:
            lens.append(cur); cur = 0
    assert_same(env, ref, "after 1000 steps")
    assert hook.steps[0] == lens and len(lens) > 10


@pytest.mark.parametrize("kind,continuous", [("CartPole", False), ("MountainCar", False), ("Pendulum", True), ("Pendulum", False)])
def test_given_actions_soft_reset_bit_exact(pkg, ctx, kind, continuous):
    n, steps = 3000, 260
    seeds = O.splitmix_states_fast(n, seed=77)
    kw = dict(continuous=continuous, n_actions=5) if kind == "Pendulum" else {}
    env = pkg.B200VecEnv(ctx, kind, n, seeds, auto_reset=False, **kw)
    params = None
    if kind == "Pendulum":
        params = O.default_params(O.KIND_PENDULUM)
        params[7], params[8] = 5, float(continuous)
    ref = O.OracleVecEnv(KINDS[kind], n, seeds, params=params)
    rng = np.random.default_rng(5)
    nact = {"CartPole": 2, "MountainCar": 3, "Pendulum": 5}[kind]
    for s in range(steps):
        if continuous:
            act = rng.uniform(-2, 2, n).astype(np.float32)
        else:
            act = rng.integers(1, nact + 1, n).astype(np.int32)
        env.reset_(is_force=False); ref.reset(force=False)
        env.act_(act); ref.step(act, auto_reset=False)
        if s % 50 == 0 or s >= steps - 2:
            assert_same(env, ref, f"{kind} step {s}")
    env.check()


@pytest.mark.parametrize("kind", ["ContinuousCartPole", "ContinuousMountainCar"])
@pytest.mark.parametrize("auto_reset", [False, True])
def test_continuous_action_variants_bit_exact(pkg, ctx, kind, auto_reset):
    """CartPoleEnv(continuous = true) (CartPoleEnv.jl:74-79,106-110) and ContinuousMountainCarEnv
    (MountainCarEnv.jl:73-74,107-111) with Float32 forces in -1.0..1.0."""
    n, steps = 4099, 420
    seeds = O.splitmix_states_fast(n, seed=311)
    env = pkg.B200VecEnv(ctx, kind, n, seeds, auto_reset=auto_reset)
    ref = O.OracleVecEnv(KINDS[kind], n, seeds)
    assert env.continuous and env.action_space() == (-1.0, 1.0)
    assert_same(env, ref, "after construction")
    rng = np.random.default_rng(17)
    for s in range(steps):
        act = rng.uniform(-1, 1, n).astype(np.float32)
        act[::7] = np.float32(1.0); act[3::11] = np.float32(-1.0)   # closed interval end points are valid
        if not auto_reset:
            env.reset_(is_force=False); ref.reset(force=False)
        env.act_(act); assert ref.step(act, auto_reset=auto_reset) == 0
        if s % 60 == 0 or s >= steps - 2:
            assert_same(env, ref, f"{kind} step {s}")
    env.check()
    assert env.episode_stats()["episodes"] > 0
    # `@assert a in action_space(env)`: 1.5 and NaN are rejected, the env is left untouched
    before = env.internal_state().copy()
    bad = np.zeros(n, np.float32); bad[5] = 1.5; bad[9] = np.nan
    env.act_(bad)
    with pytest.raises(pkg.B200RLError):
        env.check()
    after = env.internal_state()
    assert np.array_equal(bits(before[:, [5, 9]]), bits(after[:, [5, 9]]))
    with pytest.raises(pkg.B200RLError):
        env.act_random_()       # RandomPolicy over an interval is not restated
    env.close()


@pytest.mark.parametrize("kind,max_t", [("CartPole", 7), ("MountainCar", 50), ("Pendulum", 33)])
def test_max_timeout_wrapper_bit_exact(pkg, ctx, kind, max_t):
    """MaxTimeoutEnv(env, max_t) (wrappers/MaxTimeoutEnv.jl:17-28): terminal once current_t > max_t,
    reward still the wrapped env's."""
    n, steps = 2000, 180
    seeds = O.splitmix_states_fast(n, seed=5)
    kw = dict(continuous=False, n_actions=3) if kind == "Pendulum" else {}
    env = pkg.B200VecEnv(ctx, kind, n, seeds, auto_reset=True, **kw)
    params = None
    if kind == "Pendulum":
        params = O.default_params(O.KIND_PENDULUM); params[7], params[8] = 3, 0.0
    ref = O.OracleVecEnv(KINDS[kind], n, seeds, params=params)
    env.set_max_timeout(max_t); ref.set_max_timeout(max_t)
    saw_timeout = False
    for s in range(steps):
        env.act_random_(); ref.step_random(auto_reset=True)
        assert_same(env, ref, f"{kind} step {s}")
        if s == max_t - 1:
            term = env.is_terminated()
            if kind == "CartPole":      # at the timeout step the wrapped env still pays 1 where it is not itself done
                assert term.all() and (env.reward() == 1).any()
            saw_timeout = term.all()
    assert saw_timeout
    cp = env.copy(); cp.act_random_(); ref.step_random(auto_reset=True)
    assert_same(cp, ref, "copy keeps the wrapper")
    env.close(); cp.close()


def test_device_resident_actions_and_auto_reset_given_actions(pkg, ctx):
    n = 5000
    seeds = O.splitmix_states_fast(n, seed=9)
    env = pkg.B200VecEnv(ctx, "CartPole", n, seeds, auto_reset=True)
    ref = O.OracleVecEnv(O.KIND_CARTPOLE, n, seeds)
    d_act = ctx.malloc(n * 4)
    rng = np.random.default_rng(1)
    for s in range(250):
        act = rng.integers(1, 3, n).astype(np.int32)
        ctx.h2d(d_act, act)
        env.act_(d_act)
        ref.step(act, auto_reset=True)
    assert_same(env, ref, "device actions")
    ctx.free(d_act)


def test_invalid_action_sets_error_flag(pkg, ctx):
    n = 64
    env = pkg.B200VecEnv(ctx, "CartPole", n, O.splitmix_states_fast(n, 3))
    before = env.internal_state().copy()
    act = np.ones(n, np.int32); act[7] = 3
    env.act_(act)
    with pytest.raises(pkg.B200RLError) as ei:
        env.check()
    assert ei.value.status == -4
    after = env.internal_state()
    assert np.array_equal(before[:, 7], after[:, 7])       # offending env untouched
    assert not np.array_equal(before[:, 8], after[:, 8])
    env.check()  # flag cleared


def test_copy_and_seed_determinism(pkg, ctx):
    """RLBase.test_interfaces!: copies seeded identically stay == step for step (base.jl:86-118)."""
    n = 512
    env = pkg.B200VecEnv(ctx, "CartPole", n, O.splitmix_states_fast(n, 1), auto_reset=True)
    for _ in range(10):
        env.act_random_()
    c = env.copy()
    s888 = O.splitmix_states_fast(n, 888)
    env.seed_(s888); c.seed_(s888)
    for _ in range(300):
        env.act_random_(); c.act_random_()
    assert np.array_equal(bits(env.internal_state()), bits(c.internal_state()))
    assert np.array_equal(env.rng_state(), c.rng_state())
    # states stay inside state_space (base.jl:120-152)
    st = env.internal_state()
    (xlo, xhi), _, (tlo, thi), _ = env.state_space()
    assert np.all(st[0] >= xlo) and np.all(st[0] <= xhi) and np.all(st[2] >= tlo) and np.all(st[2] <= thi)


def test_episode_stats_match_host_hooks(pkg, ctx):
    n = 2048
    env = pkg.B200VecEnv(ctx, "CartPole", n, O.splitmix_states_fast(n, 21), auto_reset=True)
    steps_hook, rew_hook, dev = pkg.BatchStepsPerEpisode(n), pkg.TotalBatchRewardPerEpisode(n), pkg.DeviceEpisodeStats()
    pkg.run(pkg.RandomPolicy(), env, pkg.StopAfterNSteps(200), steps_hook + rew_hook + dev)
    n_eps = sum(len(s) for s in steps_hook.steps)
    assert dev.stats["episodes"] == n_eps and n_eps > n
    assert dev.stats["length_sum"] == sum(sum(s) for s in steps_hook.steps)
    assert dev.stats["return_sum"] == pytest.approx(sum(sum(r) for r in rew_hook.rewards))
    assert dev.stats["env_steps"] == 200 * n


@pytest.mark.parametrize("kind,continuous", [("Pendulum", True), ("Pendulum", False), ("MountainCar", False), ("ContinuousMountainCar", True)])
@pytest.mark.parametrize("auto_reset", [False, True])
def test_float64_pendulum_and_mountaincar_bit_exact(pkg, ctx, kind, continuous, auto_reset):
    """T = Float64 is what a literal `PendulumEnv()` / `MountainCarEnv()` constructs (PendulumEnv.jl:42, MountainCarEnv.jl:67): all-double
    dynamics, Float64 rand / sin / cos (the rem_pio2 decision tree), Float64 actions for the continuous spaces."""
    n, steps = 2500, 450                                   # > 2 episodes of 200 steps: theta runs up to ~ +-80 rad without wrapping
    seeds = O.splitmix_states_fast(n, seed=1234)
    kw = dict(continuous=continuous, n_actions=5) if kind == "Pendulum" else {}
    env = pkg.B200VecEnv(ctx, kind, n, seeds, T=np.float64, auto_reset=auto_reset, **kw)
    params = None
    if kind == "Pendulum":
        params = O.default_params(O.KIND_PENDULUM, "f64")
        params[7], params[8] = 5, float(continuous)
    ref = O.OracleVecEnv(KINDS[kind], n, seeds, dtype="f64", params=params)
    assert env.state().dtype == np.float64 and env.reward().dtype == np.float64
    assert_same(env, ref, "after construction")
    rng = np.random.default_rng(5)
    bound = 2.0 if kind == "Pendulum" else 1.0
    nact = {"Pendulum": 5, "MountainCar": 3}.get(kind)
    for s in range(steps):
        act = rng.uniform(-bound, bound, n) if continuous else rng.integers(1, nact + 1, n).astype(np.int32)
        if not auto_reset:
            env.reset_(is_force=False); ref.reset(force=False)
        env.act_(act); assert ref.step(act, auto_reset=auto_reset) == 0
        if s % 75 == 0 or s >= steps - 2:
            assert_same(env, ref, f"{kind} f64 step {s}")
    env.check()
    assert env.episode_stats()["episodes"] >= 2 * n
    assert env.last_action().dtype == (np.float64 if continuous else np.int32)
    env.close()


def test_float64_env_is_refused_by_the_float32_learners(pkg, ctx):
    env = pkg.B200VecEnv(ctx, "Pendulum", 64, O.splitmix_states_fast(64, 1), T=np.float64, auto_reset=True)
    desc = O.ac_desc(3, 64, 1, 0, True)
    net = pkg.Network(ctx, 3, 64, 1, O.glorot_params(desc, 1), kind=pkg.KIND_GAUSSIAN)
    with pytest.raises(pkg.B200RLError) as ei:
        pkg.OnPolicyAgent(ctx, net, env, pkg.onpolicy_config(update_freq=4, n_epochs=1, n_microbatches=1), O.splitmix_states_fast(64, 2))
    assert ei.value.status == pkg._lib.ERR_UNSUPPORTED and "Float32" in str(ei.value)
    net.close(); env.close()
