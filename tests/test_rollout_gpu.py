"""The fused rollout kernel (fwd_tc.cu: policy inference + env step + transition push for a whole stretch of steps in one
launch) against the step-by-step stage protocol (plan! -> act! -> push!, one launch each): every rollout tensor, the env
state, both RNG streams and the parameters after the update must be bit-identical, for every env / head combination."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _make(pkg, ctx, kind, n, T, seed, algo, **envkw):
    env = pkg.B200VecEnv(ctx, kind, n, O.splitmix_states_fast(n, seed), auto_reset=True, **envkw)
    n_in = {"CartPole": 4, "ContinuousCartPole": 4, "Pendulum": 3, "MountainCar": 2, "ContinuousMountainCar": 2}[kind]
    if env.continuous:
        n_out, net_kind = 1, pkg.KIND_GAUSSIAN
    else:
        n_out, net_kind = len(env.action_space()), pkg.KIND_CATEGORICAL
    desc = O.ac_desc(n_in, 64, n_out, 0, env.continuous)
    params = O.glorot_params(desc, 77)
    params = params + 0.05 * np.random.default_rng(3).standard_normal(params.size).astype(np.float32)
    net = pkg.Network(ctx, n_in, 64, n_out, params, act=0, kind=net_kind)
    cfg = pkg.onpolicy_config(update_freq=T, n_epochs=2, n_microbatches=2, algo=algo)
    agent = pkg.OnPolicyAgent(ctx, net, env, cfg, O.splitmix_states_fast(n, seed + 1), host_actions=False)
    env.reset_(is_force=True)
    return env, net, agent


CASES = [("CartPole", {}, "ppo", 1000), ("Pendulum", dict(continuous=True), "a2c", 700), ("Pendulum", dict(continuous=False, n_actions=3), "ppo", 300),
         ("MountainCar", {}, "ppo", 513), ("ContinuousCartPole", {}, "ppo", 640), ("ContinuousMountainCar", {}, "a2c", 129),
         ("CartPole", {}, "ppo", 40000)]   # 313 tiles: two resident tiles per CTA


@pytest.mark.parametrize("kind,envkw,algo,n", CASES)
def test_fused_rollout_is_bit_identical_to_the_stage_protocol(pkg, ctx, kind, envkw, algo, n):
    T = 8
    R = pkg.learners
    outs = []
    for fused in (True, False):
        env, net, agent = _make(pkg, ctx, kind, n, T, 9, algo, **envkw)
        if kind == "CartPole" and n == 1000:
            env.set_max_timeout(23)       # MaxTimeoutEnv wrapper inside the fused kernel too
        for it in range(3):               # three rollouts + updates: episodes end, envs auto-reset, parameters move
            if fused:
                agent.collect(3); agent.collect(T - 3)        # a rollout may be filled in several stretches
            else:
                for _ in range(T):
                    agent.plan(env); agent.act_fused(env); agent.push(pkg.core.PostActStage, env)
            roll = {f: agent.rollout(f) for f in (R.ROLL_ACTION, R.ROLL_LOGP, R.ROLL_REWARD, R.ROLL_TERMINAL, R.ROLL_RNG)}
            values_T = agent.rollout(R.ROLL_VALUE)[:, :T].copy()
            stats = agent.update(want_stats=True)
            roll[R.ROLL_VALUE] = agent.rollout(R.ROLL_VALUE)      # column T (bootstrap state / value) is final only after update()
            roll[R.ROLL_STATE] = agent.rollout(R.ROLL_STATE)
            assert np.array_equal(roll[R.ROLL_VALUE][:, :T], values_T)
        outs.append(dict(roll=roll, params=net.get(), state=env.internal_state(), obs=env.state(), t=env.t(), flags=env.flags(),
                         erng=env.rng_state(), rew=env.reward(), act=env.last_action(), stats=stats, ep=env.episode_stats()))
        agent.close(); net.close(); env.close()
    a, b = outs

    def same(x, y):   # bitwise, whatever the memory order (rollout tensors are Fortran-ordered)
        x, y = np.asarray(x), np.asarray(y)
        return x.shape == y.shape and x.dtype == y.dtype and x.tobytes(order="A") == y.tobytes(order="A")
    for f in a["roll"]:
        assert same(a["roll"][f], b["roll"][f]), f"rollout field {f}"
    for k in ("params", "state", "obs", "t", "flags", "erng", "rew", "act", "stats"):
        assert same(a[k], b[k]), k
    assert a["ep"]["episodes"] == b["ep"]["episodes"] and a["ep"]["length_sum"] == b["ep"]["length_sum"]
    assert a["ep"]["env_steps"] == b["ep"]["env_steps"] == 3 * T * n
    assert abs(a["ep"]["return_sum"] - b["ep"]["return_sum"]) <= 1e-6 * max(1.0, abs(b["ep"]["return_sum"]))
    if kind in ("CartPole", "ContinuousCartPole"):
        assert a["ep"]["episodes"] > 0


@pytest.mark.parametrize("kind,envkw,algo,n", [("CartPole", {}, "ppo", 1000), ("Pendulum", dict(continuous=True), "a2c", 700)])
def test_graph_replayed_iterations_are_bit_identical_to_eager_ones(pkg, ctx, kind, envkw, algo, n):
    """b200rl_onpolicy_iterate: the first iteration runs eagerly, the second is captured, the rest are graph replays — the update
    counter that keys the minibatch permutation, the grid barrier of the optimiser kernel and the Adam beta^t all live in device
    memory, so five replayed iterations must equal five collect() + update() pairs bit for bit."""
    T, iters = 8, 5
    R = pkg.learners
    outs = []
    for graph in (True, False):
        env, net, agent = _make(pkg, ctx, kind, n, T, 11, algo, **envkw)
        if graph:
            agent.iterate(2)
            stats = agent.iterate(iters - 2, want_stats=True)
            assert agent.graph_active()
        else:
            for _ in range(iters):
                agent.collect(T)
                stats = agent.update(want_stats=True)
        outs.append(dict(params=net.get(), m=net.get(R.NET_M), v=net.get(R.NET_V), bt=net.get(R.NET_BETA_T), state=env.internal_state(), erng=env.rng_state(),
                         prng=agent.rollout(R.ROLL_RNG), adv=agent.rollout(R.ROLL_ADV), stats=stats, ep=env.episode_stats(), launches=ctx.launch_count()))
        agent.close(); net.close(); env.close()
    a, b = outs
    for k in ("params", "m", "v", "bt", "state", "erng", "prng", "adv", "stats"):
        assert np.asarray(a[k]).tobytes(order="A") == np.asarray(b[k]).tobytes(order="A"), k
    assert a["ep"]["env_steps"] == b["ep"]["env_steps"] == iters * T * n and a["ep"]["episodes"] == b["ep"]["episodes"]


@pytest.mark.parametrize("kind,envkw,algo,n", [("CartPole", {}, "ppo", 4096), ("Pendulum", dict(continuous=True), "a2c", 1500)])
def test_optimiser_step_fused_into_the_loss_kernel_matches_the_separate_kernel(pkg, ctx, kind, envkw, algo, n):
    """The tensor-core loss + backward launch runs reduce -> clip_by_global_norm! -> Adam in its own tail (two in-kernel grid
    barriers); b200rl_set_fused_step(0) leaves that to reduce_clip_adam_kernel.  Both add the per-CTA gradient partials in the
    same (CTA) order, so gradients, Adam state and parameters agree bit for bit; only the FP64 sum of squares behind the
    global norm is partitioned differently (62 vs 256 parameters per CTA), which can move the norm by one ulp at most."""
    # (a continuous-action run is compared after ONE update: last-bit different parameters change the sampled Float32 actions
    #  and with them the trajectories of the following rollouts)
    T, iters = 8, (3 if kind == "CartPole" else 1)
    R = pkg.learners
    outs = []
    try:
        for fused in (1, 0):
            pkg._lib.check(ctx.lib.b200rl_set_fused_step(fused))
            env, net, agent = _make(pkg, ctx, kind, n, T, 23, algo, **envkw)
            l0 = ctx.launch_count()
            for _ in range(iters):
                agent.collect(T)
                stats = agent.update(want_stats=True)
            outs.append(dict(params=net.get(), m=net.get(R.NET_M), v=net.get(R.NET_V), bt=net.get(R.NET_BETA_T), stats=np.asarray(stats), launches=ctx.launch_count() - l0))
            agent.close(); net.close(); env.close()
    finally:
        pkg._lib.check(ctx.lib.b200rl_set_fused_step(1))
    a, b = outs
    assert a["launches"] < b["launches"], "the fused step must save one launch per optimiser step"
    assert np.array_equal(a["bt"], b["bt"])
    np.testing.assert_allclose(a["stats"], b["stats"], rtol=1e-5, atol=1e-7)   # (later iterations see last-bit different parameters)
    for k in ("params", "m", "v"):   # (the two kernels may contract Adam's multiply-adds differently: last-bit differences)
        np.testing.assert_allclose(a[k], b[k], rtol=2e-6, atol=2e-8, err_msg=k)
