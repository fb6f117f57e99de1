"""GPU test of the sharded optimiser step: two ranks in ONE process (two ctx on cuda:0, driven by two
threads) exchange gradients through the NVLink peer-exchange path of the fused reduce + clip + Adam kernel.
Checked: (1) the small all-reduce, (2) both replicas end bit-identical, (3) they match a single-rank run over
the union of the shards (same global minibatches through host permutations) to fp32 summation-order noise.
The real two-process / two-GPU wiring (CUDA IPC handles) is exercised by `bench.py --gpus 2`."""
import ctypes as C
import threading

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _in_threads(fns):
    errs = []

    def wrap(f):
        try:
            f()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=wrap, args=(f,)) for f in fns]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert not any(t.is_alive() for t in th), "a rank is stuck in the exchange"
    if errs:
        raise errs[0]


def _two_ranks(pkg):
    L = pkg._lib
    ctxs = [pkg.Context(0), pkg.Context(0)]
    regions = (C.c_void_p * 2)()
    for r, ctx in enumerate(ctxs):
        L.check(ctx.lib.b200rl_comm_init(ctx.h, 2, r, None))        # no NCCL: peer exchange only
        p = C.c_void_p()
        L.check(ctx.lib.b200rl_comm_p2p_export(ctx.h, None, C.byref(p)))
        regions[r] = p
    for ctx in ctxs:
        L.check(ctx.lib.b200rl_comm_p2p_attach(ctx.h, regions))
    return ctxs


def test_peer_exchange_small_allreduce_and_sharded_ppo_update(pkg):
    L = pkg._lib
    ctxs = _two_ranks(pkg)
    # ---- (1) small fp32 all-reduce, several rounds (slot reuse) ------------------------------------------------
    rng = np.random.default_rng(0)
    for rnd in range(5):
        host = [rng.standard_normal(700).astype(np.float32) for _ in range(2)]
        dev = [ctx.malloc(700 * 4) for ctx in ctxs]
        out = [np.empty(700, np.float32) for _ in range(2)]

        def run(r):
            ctx = ctxs[r]
            ctx.h2d(dev[r], host[r])
            L.check(ctx.lib.b200rl_comm_allreduce_f32(ctx.h, C.c_void_p(dev[r]), 700))
            ctx.d2h(out[r], dev[r])
        _in_threads([lambda r=r: run(r) for r in range(2)])
        assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], host[0] + host[1])
        [ctx.free(d) for ctx, d in zip(ctxs, dev)]

    # ---- (2)+(3) sharded PPO update vs single rank ---------------------------------------------------------------
    n_total, T, E, M = 2048, 8, 2, 2
    sh = pkg.sharding
    cfg = pkg.onpolicy_config(update_freq=T, n_epochs=E, n_microbatches=M)
    params0 = sh.glorot_actor_critic(5, 4, 64, 2)

    def make(ctx, lo, hi):
        env = pkg.B200VecEnv(ctx, "CartPole", hi - lo, sh.splitmix_states(11, lo, hi), auto_reset=True)
        net = pkg.Network(ctx, 4, 64, 2, params0.copy())
        agent = pkg.OnPolicyAgent(ctx, net, env, cfg, sh.splitmix_states(12, lo, hi))
        env.reset_(is_force=True)
        return env, net, agent
    half = n_total // 2
    nt_l = half * T
    perms_l = [np.stack([np.random.default_rng(100 * r + e).permutation(nt_l) for e in range(E)]).astype(np.int32) for r in range(2)]
    shards = [make(ctxs[r], *sh.shard_range(n_total, r, 2)) for r in range(2)]
    stats = [None, None]

    def run_rank(r):
        env, net, agent = shards[r]
        for it in range(2):
            agent.collect(T)
            stats[r] = agent.update(perms_l[r], want_stats=True)
    _in_threads([lambda r=r: run_rank(r) for r in range(2)])
    p = [shards[r][1].get() for r in range(2)]
    assert np.array_equal(p[0], p[1]), "replicas diverged"                       # (2)
    assert np.array_equal(stats[0], stats[1]) and np.isfinite(stats[0]).all()
    assert not np.array_equal(p[0], params0)

    # single rank over all envs; global minibatch k = union of the shards' minibatch k (flat index = env + N * t)
    ctx1 = pkg.Context(0)
    env1, net1, agent1 = make(ctx1, 0, n_total)
    B_l = nt_l // M
    perm_g = np.empty((E, n_total * T), np.int32)
    for e in range(E):
        for mb in range(M):
            parts = []
            for r in range(2):
                li = perms_l[r][e, mb * B_l:(mb + 1) * B_l]
                env_l, t = li % half, li // half
                parts.append((env_l + r * half) + n_total * t)
            perm_g[e, mb * 2 * B_l:(mb + 1) * 2 * B_l] = np.concatenate(parts)
    for it in range(2):
        agent1.collect(T)
        st1 = agent1.update(perm_g, want_stats=True)
    p1 = net1.get()
    # same samples, same global normalisation; only the fp32 summation order differs (per-CTA partials, rank order)
    np.testing.assert_allclose(stats[0][:, :5], st1[:, :5], rtol=2e-4, atol=2e-6)
    diff = np.abs(p[0] - p1)
    assert np.mean(diff <= 2e-5) > 0.995 and diff.max() < 3e-4                   # (3)
    for env, net, agent in shards + [(env1, net1, agent1)]:
        agent.close(); net.close(); env.close()
