"""GPU tests of the DQN action path (BASELINE config 5 driven through run()):
b200rl_net_q_explore = BatchExplorer(EpsilonGreedyExplorer) with the per-column decay schedule, bit-exact against the
oracle restatement of epsilon_greedy_explorer.jl:69-112 / batch_explorer.jl:15-21 given the same Q table and streams;
Agent(QBasedPolicy(DQNLearner, explorer), Trajectory(prioritised ring, BatchSampler, InsertSampleRatioController)) end
to end through run(policy, env, stop, hook)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def q_explore(pkg, ctx, net, obs, rng, explorer):
    """Returns (actions, advanced streams) of one b200rl_net_q_explore call."""
    n = obs.shape[1]
    d_obs = ctx.malloc(obs.nbytes); ctx.h2d(d_obs, np.asfortranarray(obs))
    d_rng = ctx.malloc(n * 32); ctx.h2d(d_rng, rng)
    d_act = ctx.malloc(n * 4)
    st = explorer.as_struct()
    pkg._lib.check(ctx.lib.b200rl_net_q_explore(net.h, C.c_void_p(d_obs), n, C.c_void_p(d_rng), C.byref(st), C.c_void_p(d_act)))
    a = np.empty(n, np.int32); ctx.d2h(a, d_act)
    r = np.empty((n, 4), np.uint64); ctx.d2h(r, d_rng)
    for d in (d_obs, d_rng, d_act):
        ctx.free(d)
    return a, r


@pytest.mark.parametrize("kind,brk,na,hidden", [("linear", False, 2, 128), ("exp", False, 3, 64), ("linear", True, 2, 128), ("exp", True, 4, 64)])
def test_q_explore_bit_exact_given_the_q_table(pkg, ctx, kind, brk, na, hidden):
    ns, n = 4, 6000
    desc = O.ac_desc(ns, hidden, na)
    p = O.glorot_params(desc, 31, q_net=True)
    net = pkg.Network(ctx, ns, hidden, na, p, kind=pkg.KIND_Q)
    obs = np.asfortranarray(np.random.default_rng(1).standard_normal((ns, n)).astype(np.float32))
    qv = net.values(obs)                                      # the device's own Q table: selection parity is then exact
    # the batch straddles warm-up (eps_init), the decay and (linear) the stable tail
    ex = pkg.EpsilonGreedyExplorer(0.05, kind=kind, eps_init=0.9, warmup_steps=1500, decay_steps=3000, step=500, is_break_tie=brk)
    ex6 = O.explorer6(0.05, 0.9, 1500, 3000, kind, brk)
    seeds = O.splitmix_states_fast(n, 99)
    a, r = q_explore(pkg, ctx, net, obs, seeds, ex)
    ref_rng = seeds.copy()
    ref = O.egreedy_plan(ex6, 500, qv, ref_rng)
    assert np.array_equal(a, ref)
    assert np.array_equal(r, ref_rng)
    greedy = qv.argmax(0) + 1
    assert (a[:1000] == greedy[:1000]).mean() < (a[-1000:] == greedy[-1000:]).mean()   # eps decays across the batch
    net.close()


def test_q_explore_break_tie_on_an_all_tie_table_and_zero_epsilon(pkg, ctx):
    ns, na, n = 4, 3, 30000
    desc = O.ac_desc(ns, 64, na)
    net = pkg.Network(ctx, ns, 64, na, np.zeros(O.q_nparams(desc), np.float32), kind=pkg.KIND_Q)   # Q == 0 everywhere: every action ties
    obs = np.asfortranarray(np.random.default_rng(2).standard_normal((ns, n)).astype(np.float32))
    qv = np.zeros((na, n), np.float32, order="F")
    seeds = O.splitmix_states_fast(n, 5)
    for brk in (True, False):
        ex = pkg.EpsilonGreedyExplorer(0.0, eps_init=0.0, is_break_tie=brk)      # eps = 0: the uniform is still drawn
        a, r = q_explore(pkg, ctx, net, obs, seeds, ex)
        ref_rng = seeds.copy()
        ref = O.egreedy_plan(O.explorer6(0.0, 0.0, 0, 0, "linear", brk), 1, qv, ref_rng)
        assert np.array_equal(a, ref) and np.array_equal(r, ref_rng) and not np.array_equal(r, seeds)
        if brk:
            np.testing.assert_allclose(np.bincount(a, minlength=na + 1)[1:] / n, 1.0 / na, atol=0.01)
        else:
            assert np.all(a == 1)                                                 # findmax: the first maximum
    net.close()


def test_q_explore_rejects_bad_arguments(pkg, ctx):
    desc = O.ac_desc(4, 64, 2)
    net = pkg.Network(ctx, 4, 64, 2, O.glorot_params(desc, 1, q_net=True), kind=pkg.KIND_Q)
    d = ctx.malloc(1024)
    st = pkg.EpsilonGreedyExplorer(1.5).as_struct()
    assert ctx.lib.b200rl_net_q_explore(net.h, C.c_void_p(d), 8, C.c_void_p(d), C.byref(st), C.c_void_p(d)) == pkg._lib.ERR_INVALID
    st = pkg.EpsilonGreedyExplorer(0.1).as_struct()
    assert ctx.lib.b200rl_net_q_explore(net.h, C.c_void_p(d), 8, None, C.byref(st), C.c_void_p(d)) == pkg._lib.ERR_INVALID
    ac = pkg.Network(ctx, 4, 64, 2, O.glorot_params(desc, 1), kind=pkg.KIND_CATEGORICAL)
    assert ctx.lib.b200rl_net_q_explore(ac.h, C.c_void_p(d), 8, C.c_void_p(d), C.byref(st), C.c_void_p(d)) == pkg._lib.ERR_INVALID
    ctx.free(d); net.close(); ac.close()


@pytest.mark.parametrize("host_actions", [False, True])
def test_dqn_agent_through_run(pkg, ctx, host_actions):
    """Config-5 shape at test size: prioritised ring, eps-greedy exploration with a decay schedule, one update per
    inserted frame after a threshold, target sync — driven by run(agent, env, StopAfterNSteps, hook)."""
    lanes, cap, B, steps, hidden = 256, 64, 512, 48, 128
    env_seeds = O.splitmix_states_fast(lanes, 21)
    env = pkg.B200VecEnv(ctx, "CartPole", lanes, env_seeds, auto_reset=True)
    desc = O.ac_desc(4, hidden, 2)
    p0 = O.glorot_params(desc, 8, q_net=True)
    net = pkg.Network(ctx, 4, hidden, 2, p0.copy(), kind=pkg.KIND_Q)
    traj = pkg.Trajectory(ctx, 4, cap, lanes=lanes, batch_size=B, sampler_rng=O.splitmix_states_fast(B, 22), prioritized=True)
    traj.controller = pkg.InsertSampleRatioController(ratio=1.0, threshold=8)
    learner = pkg.DQNLearner(ctx, net, traj, pkg.dqn_config(target_update_freq=10))
    explorer = pkg.EpsilonGreedyExplorer(0.01, kind="exp", eps_init=1.0, warmup_steps=2 * lanes, decay_steps=10 * lanes)
    ex_seeds = O.splitmix_states_fast(lanes, 23)
    policy = pkg.QBasedPolicy(ctx, learner, explorer, ex_seeds, lanes)
    agent = pkg.Agent(policy, traj, host_actions=host_actions)

    # the first planned batch, before any update: same streams + schedule as the oracle on the device's Q table
    probe = env.copy(); probe.reset_(is_force=True)
    qv0 = net.values(probe.state())
    ref_rng = ex_seeds.copy()
    a0 = O.egreedy_plan(O.explorer6(0.01, 1.0, 2 * lanes, 10 * lanes, "exp"), 1, qv0, ref_rng)
    probe.close()

    first = {}

    class Spy(pkg.AbstractHook):
        def push(self, stage, policy_, env_):
            if stage == "PostActStage" and "a" not in first:
                first["a"] = env_.last_action(); first["rng"] = policy.explorer_rng()

    hook = pkg.run(agent, env, pkg.StopAfterNSteps(steps), Spy() + pkg.BatchStepsPerEpisode(lanes))
    assert np.array_equal(first["a"], a0) and np.array_equal(first["rng"], ref_rng)
    n_term = np.array([len(e) for e in hook[1][()]])
    assert np.array_equal(traj.lane_lengths(), steps + n_term)   # steps + episodes - 1 per lane (RLCore/test/core/base.jl:20)
    assert traj.n_sampleable() == steps * lanes
    assert explorer.step == 1 + steps * lanes                   # BatchExplorer: one inner step per column
    c = traj.controller
    assert c.n_inserted == steps and c.n_sampled == steps - 8 + 1
    assert not np.array_equal(net.get(), p0)                    # the learner trained
    assert np.array_equal(net.get(pkg.learners.NET_TARGET), net.get()) == ((c.n_sampled % 10) == 0)
    done = sum(len(s) for s in hook[1][()])
    assert done == env.episode_stats()["episodes"] > 0          # host hook and device-side statistics agree
    pkg.run(agent, env, pkg.StopAfterNSteps(3))                  # re-entry on a filled ring: the forced reset starts an episode in every lane
    assert traj.n_sampleable() == min((steps + 3), cap) * lanes or traj.n_sampleable() <= cap * lanes
    assert (traj.lane_lengths() <= cap).all()
    # the ring holds what the env produced: replay the stored actions through the oracle env
    b = traj.sample(beta=0.4)
    assert set(np.unique(b["action"])) <= {1, 2} and np.isfinite(b["state"]).all()
    policy.close(); traj.close(); net.close(); env.close()


# ---- checkpoint / resume (the JLD2 hook pattern of docs/src/How_to_use_hooks.md:124-167) ---------------------
def _ppo_objects(pkg, ctx, kind, n, T, seed):
    gauss = kind == "Pendulum"
    env = pkg.B200VecEnv(ctx, kind, n, O.splitmix_states_fast(n, seed), auto_reset=True)
    n_in, n_out = (3, 1) if gauss else (4, 2)
    desc = O.ac_desc(n_in, 64, n_out, 0, gauss)
    net = pkg.Network(ctx, n_in, 64, n_out, O.glorot_params(desc, seed + 1), kind=pkg.KIND_GAUSSIAN if gauss else pkg.KIND_CATEGORICAL)
    cfg = pkg.onpolicy_config(update_freq=T, n_epochs=2, n_microbatches=2, algo="a2c" if gauss else "ppo")
    agent = pkg.OnPolicyAgent(ctx, net, env, cfg, O.splitmix_states_fast(n, seed + 2), host_actions=False)
    return env, net, agent


def _continue(agent, T, t_now):
    """finish the current rollout, update, one more full iteration, then 3 steps into the next rollout"""
    agent.collect(T - t_now); s1 = agent.update(want_stats=True)
    agent.collect(T); s2 = agent.update(want_stats=True)
    agent.collect(3)
    return np.concatenate([s1, s2])


@pytest.mark.parametrize("kind,mid", [("CartPole", 3), ("CartPole", 0), ("Pendulum", 5)])
def test_checkpoint_resume_is_bit_identical(pkg, ctx, tmp_path, kind, mid):
    n, T = 640, 8
    env, net, agent = _ppo_objects(pkg, ctx, kind, n, T, 40)
    env.reset_(is_force=True)
    agent.collect(T); agent.update()
    agent.collect(T); agent.update()
    if mid:
        agent.collect(mid)                                               # checkpoint in the middle of a rollout
    ck = pkg.checkpoint.checkpoint(env=env, net=net, agent=agent)
    path = str(tmp_path / "ck.npz")
    pkg.checkpoint.save(path, ck)
    stats_a = _continue(agent, T, mid)
    final_a = pkg.checkpoint.checkpoint(env=env, net=net, agent=agent)

    # fresh objects with DIFFERENT seeds / parameters: everything that matters must come from the checkpoint
    env2, net2, agent2 = _ppo_objects(pkg, ctx, kind, n, T, 777)
    env2.reset_(is_force=True)
    agent2.collect(2)
    pkg.checkpoint.restore(pkg.checkpoint.load(path), env=env2, net=net2, agent=agent2)
    assert agent2.fill() == (mid, T)
    back = pkg.checkpoint.checkpoint(env=env2, net=net2, agent=agent2)
    assert sorted(back) == sorted(ck)
    for k in ck:
        assert np.array_equal(np.asarray(back[k]), np.asarray(ck[k])), k     # restore o checkpoint = identity
    stats_b = _continue(agent2, T, mid)
    final_b = pkg.checkpoint.checkpoint(env=env2, net=net2, agent=agent2)
    assert np.array_equal(stats_a, stats_b)
    for k in final_a:
        assert np.array_equal(np.asarray(final_a[k]), np.asarray(final_b[k])), k
    assert env.episode_stats() == env2.episode_stats()
    for o in (agent, agent2, net, net2, env, env2):
        o.close()


def test_checkpoint_hook_writes_files(pkg, ctx, tmp_path):
    env, net, agent = _ppo_objects(pkg, ctx, "CartPole", 256, 4, 3)
    agent.fusable = False                                                # a per-step hook steps through the stages anyway
    hook = pkg.checkpoint.CheckpointEveryNSteps(5, str(tmp_path / "run"), net=net)
    pkg.run(agent, env, pkg.StopAfterNSteps(11), hook)
    assert [p.split("_")[-1] for p in hook.paths] == ["5.npz", "10.npz"]
    ck = pkg.checkpoint.load(hook.paths[1])
    assert ck["agent/counters"].tolist()[:2] == [10 % 4, 10 // 4] and ck["env/state"].shape == (4, 256)
    agent.close(); net.close(); env.close()


def _dqn_objects(pkg, ctx, lanes, cap, B, seed):
    env = pkg.B200VecEnv(ctx, "CartPole", lanes, O.splitmix_states_fast(lanes, seed), auto_reset=True)
    desc = O.ac_desc(4, 64, 2)
    net = pkg.Network(ctx, 4, 64, 2, O.glorot_params(desc, seed + 1, q_net=True), kind=pkg.KIND_Q)
    traj = pkg.Trajectory(ctx, 4, cap, lanes=lanes, batch_size=B, sampler_rng=O.splitmix_states_fast(B, seed + 2), prioritized=True)
    traj.controller = pkg.InsertSampleRatioController(ratio=1.0, threshold=4)
    learner = pkg.DQNLearner(ctx, net, traj, pkg.dqn_config(target_update_freq=5))
    explorer = pkg.EpsilonGreedyExplorer(0.05, kind="linear", eps_init=1.0, warmup_steps=lanes, decay_steps=20 * lanes)
    policy = pkg.QBasedPolicy(ctx, learner, explorer, O.splitmix_states_fast(lanes, seed + 3), lanes)
    return env, net, traj, policy, pkg.Agent(policy, traj)


def test_dqn_checkpoint_resume_is_bit_identical(pkg, ctx):
    """The replay run's whole state — env, Q-network + target + Adam state + update counter, the ring with its per-lane bookkeeping and
    sum tree, sampler / explorer streams, explorer step, controller counters — through checkpoint_replay / restore_replay: a run that
    is stopped, rebuilt from different seeds, restored and continued must equal the uninterrupted one bit for bit."""
    lanes, cap, B = 64, 40, 128
    ck_mod = pkg.checkpoint

    def continue_(env, agent):
        env_steps = 0
        for _ in range(30):                                     # the stage loop of run() without the forced reset of a new run
            agent.push(pkg.core.PreActStage, env)
            agent.act_fused(env) if isinstance(agent.plan(env), pkg.core.FusedAction) else None
            agent.push(pkg.core.PostActStage, env)
            agent.optimise(pkg.core.PostActStage)
            env_steps += 1
        return env_steps

    env, net, traj, policy, agent = _dqn_objects(pkg, ctx, lanes, cap, B, 50)
    pkg.run(agent, env, pkg.StopAfterNSteps(25), pkg.EmptyHook())
    ck = ck_mod.checkpoint_replay(env, net, agent)
    continue_(env, agent)
    final_a = ck_mod.checkpoint_replay(env, net, agent)

    env2, net2, traj2, policy2, agent2 = _dqn_objects(pkg, ctx, lanes, cap, B, 999)
    pkg.run(agent2, env2, pkg.StopAfterNSteps(3), pkg.EmptyHook())
    ck_mod.restore_replay(ck, env2, net2, agent2)
    back = ck_mod.checkpoint_replay(env2, net2, agent2)
    assert sorted(back) == sorted(ck)
    for k in ck:
        assert np.array_equal(np.asarray(back[k]), np.asarray(ck[k])), k
    continue_(env2, agent2)
    final_b = ck_mod.checkpoint_replay(env2, net2, agent2)
    for k in final_a:
        assert np.array_equal(np.asarray(final_a[k]), np.asarray(final_b[k])), k
    for o in (policy, policy2, traj, traj2, net, net2, env, env2):
        o.close()
