"""GPU parity of K3/K4 (trajectory ring push, uniform / prioritised sampling + gather, sum-tree
priority updates) and the DQN update (BASELINE config 5 shape) against the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def fill_both(pkg, ctx, ns, lanes, cap, frames, prioritized, B, seed=0, default_priority=1.0):
    slots = O.splitmix_states_fast(B, 900 + seed)
    tr = pkg.Trajectory(ctx, ns, cap, lanes=lanes, batch_size=B, sampler_rng=slots, prioritized=prioritized, default_priority=default_priority)
    ref = O.OracleTraj(ns, lanes, cap, prioritized, default_priority)
    rng = np.random.default_rng(seed)
    obs = rng.standard_normal((ns, lanes)).astype(np.float32)
    tr.push_state(obs); ref.push_state(obs)
    assert len(tr) == len(ref) == 0
    for k in range(frames):
        a = rng.integers(1, 3, lanes).astype(np.int32); r = rng.standard_normal(lanes).astype(np.float32)
        t = (rng.random(lanes) < 0.1).astype(np.uint8); obs = rng.standard_normal((ns, lanes)).astype(np.float32)
        tr.push(a, r, t, obs); ref.push(a, r, t, obs)
        assert len(tr) == len(ref) == min(k + 1, cap)
    return tr, ref, slots


def assert_batches_equal(b, rb):
    for k in ("state", "action", "reward", "terminal", "next_state", "key"):
        assert np.array_equal(b[k], rb[k]), k


@pytest.mark.parametrize("lanes,cap,frames", [(1, 16, 40), (8, 32, 20), (64, 10, 35)])
def test_uniform_sampling_bit_exact(pkg, ctx, lanes, cap, frames):
    B = 256
    tr, ref, slots = fill_both(pkg, ctx, 4, lanes, cap, frames, False, B)
    s = slots.copy()
    for _ in range(3):
        b = tr.sample()
        rb = ref.sample(s, B)         # oracle advances `s` in place
        assert_batches_equal(b, rb)
    assert np.array_equal(tr.sampler_rng(), s)


@pytest.mark.parametrize("lanes,cap,frames", [(1, 64, 100), (16, 32, 50)])
def test_prioritized_sampling_and_updates_bit_exact(pkg, ctx, lanes, cap, frames):
    B = 512
    tr, ref, slots = fill_both(pkg, ctx, 4, lanes, cap, frames, True, B, seed=3, default_priority=2.0)
    assert tr.total_priority() == ref.total_priority() == 2.0 * cap * lanes
    s = slots.copy()
    rng = np.random.default_rng(0)
    for it in range(4):
        b = tr.sample(beta=0.5)
        rb = ref.sample(s, B, prioritized=True, beta=0.5)
        assert_batches_equal(b, rb)
        assert np.array_equal(b["priority"], rb["priority"])
        np.testing.assert_allclose(b["weight"], rb["weight"], rtol=2e-6)
        prio = (rng.random(B) ** 0.6 + 1e-3).astype(np.float32)
        for k in np.unique(b["key"]):          # duplicates in a batch must carry one value
            prio[b["key"] == k] = prio[np.argmax(b["key"] == k)]
        tr.update_priority(prio); ref.update_priority(b["key"], prio)
        assert tr.total_priority() == ref.total_priority()
    # pushing after priority updates keeps both trees in step (wrap-around drops old leaves)
    obs = np.zeros((4, lanes), np.float32)
    for _ in range(5):
        tr.push(np.ones(lanes, np.int32), np.zeros(lanes, np.float32), np.zeros(lanes, np.uint8), obs)
        ref.push(np.ones(lanes, np.int32), np.zeros(lanes, np.float32), np.zeros(lanes, np.uint8), obs)
        assert tr.total_priority() == ref.total_priority()


def _drive(pkg, ctx, env, tr, ref, steps, hook):
    """random-policy steps with trajectory pushes the way learners.Agent does them; the oracle ring gets the same calls"""
    for _ in range(steps):
        if not env.auto_reset:
            env.reset_(is_force=False)                                        # soft reset of the finished sub-envs ...
            tr.push_env(env, first_state_only=2); ref.push_episode_start(env.state(), pending_only=True)   # ... whose episodes start here
        env.act_random_()
        tr.push_env(env)
        ref.push(env.last_action(), env.reward(), env.flags(), env.state())
        hook.push("PostActStage", None, env)


@pytest.mark.parametrize("auto_reset", [True, False])
def test_episodes_buffer_length_identity_per_lane(pkg, ctx, auto_reset):
    """RLCore/test/core/base.jl:20 per lane: length(container) == steps + episodes - 1 — every episode's first state is a frame of its
    own and the entry straddling two episodes is stored but never sampleable; the sampleable entries are exactly the steps.
    (With the in-kernel auto-reset an episode that ends on the very last step has already started its successor: that frame counts.)"""
    n, steps, B = 96, 70, 2048
    env = pkg.B200VecEnv(ctx, "CartPole", n, O.splitmix_states_fast(n, 5), auto_reset=auto_reset)
    slots = O.splitmix_states_fast(B, 6)
    tr = pkg.Trajectory(ctx, 4, 200, lanes=n, batch_size=B, sampler_rng=slots)
    ref = O.OracleTraj(4, n, 200)
    hook = pkg.BatchStepsPerEpisode(n)
    env.reset_(is_force=True)
    tr.push_env(env, first_state_only=True); ref.push_state(env.state())     # PreEpisodeStage after the forced reset
    assert np.array_equal(tr.lane_lengths(), np.zeros(n, np.int64))           # test/policies/agent.jl:30 (0 after the first state)
    _drive(pkg, ctx, env, tr, ref, steps, hook)
    finished = np.array([len(s) for s in hook[()]])
    assert finished.sum() > n                                                 # CartPole under the random policy: ~3 episodes per lane
    last_step_terminal = env.is_terminated()
    started = 1 + finished - (0 if auto_reset else last_step_terminal.astype(int))   # soft reset: the successor of a just-finished episode has not begun
    lengths = tr.lane_lengths()
    assert np.array_equal(lengths, steps + started - 1)
    assert np.array_equal(lengths, ref.lane_lengths())
    assert tr.n_sampleable() == ref.n_sampleable() == steps * n               # every step is sampleable, no straddling entry is
    s = slots.copy()
    b = tr.sample(); rb = ref.sample(s, B)
    assert_batches_equal(b, rb) and np.array_equal(tr.sampler_rng(), s)
    # a sampled terminal transition really ends an episode; a non-terminal one is followed by its true successor state
    term = b["terminal"].astype(bool)
    assert term.any() and (~term).any()
    # forced reset in the middle (ResetAfterNSteps / re-entering run): an episode-start frame for every lane, nothing else changes
    env.reset_(is_force=True)
    tr.push_env(env, first_state_only=True); ref.push_state(env.state())
    assert np.array_equal(tr.lane_lengths(), lengths + 1) and tr.n_sampleable() == steps * n
    _drive(pkg, ctx, env, tr, ref, 5, hook)
    assert tr.n_sampleable() == ref.n_sampleable() == (steps + 5) * n
    assert_batches_equal(tr.sample(), ref.sample(s, B))
    tr.close(); env.close()


def test_single_lane_matches_agent_jl_lengths_and_wraps(pkg, ctx):
    """test/policies/agent.jl:27-34 (length 0 after the first state, 1 after the first transition) with lanes = 1, then far past the
    capacity: the ring keeps the newest cap entries, overwritten entries leave the sampleable set (and the sum tree)."""
    cap, B = 16, 64
    slots = O.splitmix_states_fast(B, 1)
    tr = pkg.Trajectory(ctx, 4, cap, lanes=1, batch_size=B, sampler_rng=slots, prioritized=True, default_priority=1.5)
    ref = O.OracleTraj(4, 1, cap, True, 1.5)
    rng = np.random.default_rng(0)
    obs = rng.standard_normal((4, 1)).astype(np.float32)
    tr.push_state(obs); ref.push_state(obs)
    assert len(tr) == 0
    for k in range(60):
        term = np.uint8(1 if k % 7 == 6 else 0)
        nxt = rng.standard_normal((4, 1)).astype(np.float32)
        a, r, t = np.array([1 + k % 2], np.int32), np.array([k], np.float32), np.array([term], np.uint8)
        tr.push(a, r, t, nxt); ref.push(a, r, t, nxt)
        if k == 0:
            assert len(tr) == 1
        if term:                                                    # PreEpisodeStage of the next episode (the reference's reset + push)
            s0 = rng.standard_normal((4, 1)).astype(np.float32)
            tr.push_episode_start(s0, pending_only=True); ref.push_episode_start(s0, pending_only=True)
        assert len(tr) == len(ref) <= cap and tr.n_sampleable() == ref.n_sampleable()
        assert tr.total_priority() == ref.total_priority() == 1.5 * tr.n_sampleable()
    s = slots.copy()
    b = tr.sample(beta=0.4); rb = ref.sample(s, B, prioritized=True, beta=0.4)
    assert_batches_equal(b, rb) and np.array_equal(b["priority"], rb["priority"])
    assert set(b["reward"].astype(int)) <= set(range(60 - cap - 3, 60))     # only recent transitions survive
    tr.close()


def test_push_env_matches_host_push(pkg, ctx):
    n = 128
    env = pkg.B200VecEnv(ctx, "CartPole", n, O.splitmix_states_fast(n, 12), auto_reset=True)
    slots = O.splitmix_states_fast(64, 1)
    tr = pkg.Trajectory(ctx, 4, 50, lanes=n, batch_size=64, sampler_rng=slots)
    ref = O.OracleTraj(4, n, 50)
    env.reset_(is_force=True)
    tr.push_env(env, first_state_only=True); ref.push_state(env.state())
    for _ in range(30):
        env.act_random_()
        tr.push_env(env)
        ref.push(env.last_action(), env.reward(), env.flags(), env.state())     # flags: bit1 = the env has already auto-reset
    assert np.array_equal(tr.lane_lengths(), ref.lane_lengths()) and tr.n_sampleable() == ref.n_sampleable() == 30 * n
    s = slots.copy()
    assert_batches_equal(tr.sample(), ref.sample(s, 64))


@pytest.mark.parametrize("hidden,huber,double_dqn,prioritized", [(128, True, False, True), (64, False, False, False), (128, True, True, True)])
def test_dqn_update_parity(pkg, ctx, hidden, huber, double_dqn, prioritized):
    """Config-5 shaped: Q-net 4 -> hidden -> hidden -> 2, prioritised replay, sample + TD update + priority write-back."""
    ns, na, B, lanes = 4, 2, 1024, 32
    tr, ref, slots = fill_both(pkg, ctx, ns, lanes, 64, 80, prioritized, B, seed=5)
    desc = O.ac_desc(ns, hidden, na)
    p0 = O.glorot_params(desc, 2, q_net=True) + 0.05 * np.random.default_rng(1).standard_normal(O.q_nparams(desc)).astype(np.float32)
    net = pkg.Network(ctx, ns, hidden, na, p0, kind=pkg.KIND_Q)
    net.set(pkg.learners.NET_TARGET, p0 * np.float32(0.9))
    cfg = pkg.dqn_config(huber=huber, double_dqn=double_dqn, target_update_freq=3, max_grad_norm=10.0, per_beta=0.4)
    learner = pkg.DQNLearner(ctx, net, tr, cfg)
    p = p0.copy(); pt = p0 * np.float32(0.9)
    m = np.zeros_like(p); v = np.zeros_like(p); bt = np.array([0.9, 0.999], np.float32)
    q_ref = O.q_values(desc, p, np.asfortranarray(np.random.default_rng(3).standard_normal((ns, 100)).astype(np.float32)))
    np.testing.assert_allclose(net.values(np.random.default_rng(3).standard_normal((ns, 100)).astype(np.float32)), q_ref, rtol=1e-5, atol=2e-6)
    for it in range(4):
        stats = learner.update(want_stats=True)
        b = tr.batch()                                       # the batch the update used (teacher forcing for the oracle)
        w = b["weight"] if prioritized else None
        g, loss, td = O.dqn_loss_grad(desc, p, pt, b["state"], b["action"], b["reward"], b["terminal"], b["next_state"], w, 0.99, huber, double_dqn)
        gc, gn = O.clip_by_global_norm(g.astype(np.float32), 10.0)
        O.adam_step(p, gc, m, v, bt)
        tol = 2e-5 * (1 + it)
        assert stats["loss"] == pytest.approx(loss, rel=tol)
        assert stats["grad_norm"] == pytest.approx(gn, rel=10 * tol)
        np.testing.assert_allclose(learner.last_td(), td, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(net.get(), p, rtol=0, atol=5e-6)
        if prioritized:
            newp = (np.abs(td) + np.float32(1e-6)) ** np.float32(0.6)
            ref.update_priority(b["key"], newp.astype(np.float32))
            assert tr.total_priority() == pytest.approx(ref.total_priority(), rel=1e-5)
        if (it + 1) % 3 == 0:
            pt = p.copy()                                    # hard target sync (rho = 0)
            np.testing.assert_allclose(net.get(pkg.learners.NET_TARGET), net.get(), rtol=0, atol=0)
        else:
            assert not np.array_equal(net.get(pkg.learners.NET_TARGET), net.get())


def test_q_act_epsilon_greedy(pkg, ctx):
    import ctypes as C
    ns, na, n = 4, 2, 20000
    desc = O.ac_desc(ns, 128, na)
    p = O.glorot_params(desc, 4, q_net=True)
    net = pkg.Network(ctx, ns, 128, na, p, kind=pkg.KIND_Q)
    obs = np.asfortranarray(np.random.default_rng(0).standard_normal((ns, n)).astype(np.float32))
    d_obs = ctx.malloc(obs.nbytes); ctx.h2d(d_obs, obs)
    d_rng = ctx.malloc(n * 32); ctx.h2d(d_rng, O.splitmix_states_fast(n, 8))
    d_act = ctx.malloc(n * 4)
    greedy = O.q_values(desc, p, obs).argmax(0) + 1
    for eps, lo, hi in ((0.0, 1.0, 1.0), (0.5, 0.70, 0.80)):
        pkg._lib.check(ctx.lib.b200rl_net_q_act(net.h, C.c_void_p(d_obs), n, C.c_void_p(d_rng), C.c_float(eps), C.c_void_p(d_act)))
        a = np.empty(n, np.int32); ctx.d2h(a, d_act)
        agree = np.mean(a == greedy)
        assert lo - 0.01 <= agree <= hi + 0.01 and a.min() >= 1 and a.max() <= na
    for d in (d_obs, d_rng, d_act):
        ctx.free(d)
