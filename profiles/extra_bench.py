"""Secondary measurements (not the bench.py headline): BASELINE configs 3 and 5 and the
bandwidth-saturating sweep of the env kernel.  Run under gpurun; prints JSON lines."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
b = g.load_package()
sh = b.sharding
ctx = b.Context(0)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]

def timed(fn, reps, warm=3):
    for _ in range(warm): fn()
    ctx.sync(); ctx.timer_start()
    for _ in range(reps): fn()
    return ctx.timer_stop_ms() / reps

# ---- K1 sweep: random-policy CartPole step (state+rng every step) and policy-action step -------
for n in (65536, 1 << 20, 1 << 22, 1 << 24):
    env = b.B200VecEnv(ctx, "CartPole", n, sh.splitmix_states(1, 0, n), auto_reset=True)
    env.reset_(True)
    ms_r = timed(env.act_random_, 50)
    d_act = ctx.malloc(n * 4); ctx.h2d(d_act, np.ones(n, np.int32))
    ms_a = timed(lambda: env.act_(d_act), 50)
    # algorithmic bytes: 57 B/env-step given actions (+64 B rng R/W per step with the fused random policy)
    print(json.dumps({"bench": "K1 env_step sweep", "n_envs": n, "ms_random_policy": ms_r, "ms_given_actions": ms_a,
                      "steps_per_s_given": n / ms_a * 1e3, "gbs_given": n * 57 / ms_a / 1e6, "frac_hbm_given": n * 57 / ms_a / 1e6 / HBM,
                      "gbs_random": n * (57 + 64 - 4) / ms_r / 1e6, "frac_hbm_random": n * (57 + 64 - 4) / ms_r / 1e6 / HBM}), flush=True)
    ctx.free(d_act); env.close()

# ---- K5 sweep ---------------------------------------------------------------------------------
import ctypes as C
for n, T in ((65536, 32), (1 << 20, 32), (1 << 22, 32)):
    nb = n * T
    d = {k: ctx.malloc(nb * 4 + n * 4) for k in ("r", "v", "adv")}
    dt = ctx.malloc(nb)
    b._lib.check(ctx.lib.b200rl_memset(ctx.h, C.c_void_p(d["r"]), 0, nb * 4)); b._lib.check(ctx.lib.b200rl_memset(ctx.h, C.c_void_p(d["v"]), 0, nb * 4 + n * 4))
    b._lib.check(ctx.lib.b200rl_memset(ctx.h, C.c_void_p(dt), 0, nb))
    f = lambda: b._lib.check(ctx.lib.b200rl_gae_f32(ctx.h, C.c_void_p(d["adv"]), C.c_void_p(d["r"]), C.c_void_p(d["v"]), C.c_void_p(dt), C.c_float(0.99), C.c_float(0.95), n, T, 2, 1))
    ms = timed(f, 30)
    print(json.dumps({"bench": "K5 GAE sweep (adv only, 13 B/elt)", "n_series": n, "T": T, "ms": ms, "gbs": nb * 13 / ms / 1e6, "frac_hbm": nb * 13 / ms / 1e6 / HBM}), flush=True)
    for k in d: ctx.free(d[k])
    ctx.free(dt)

# ---- config 3: 32768 Pendulum + A2C (GAE lambda .95), continuous Gaussian head -------------------
n, T = 32768, 32
env = b.B200VecEnv(ctx, "Pendulum", n, sh.splitmix_states(3, 0, n), auto_reset=True)
r = np.random.default_rng(5)
def dense(o, i):
    lim = np.sqrt(6.0 / (i + o)); return [r.uniform(-lim, lim, (o, i)).astype(np.float32).ravel(order="F"), np.zeros(o, np.float32)]
params = np.concatenate(dense(64, 3) + dense(64, 64) + dense(1, 64) + dense(1, 64) + dense(64, 3) + dense(64, 64) + dense(1, 64))
net = b.Network(ctx, 3, 64, 1, params, act=b.ACT_TANH, kind=b.KIND_GAUSSIAN)
cfg = b.onpolicy_config(update_freq=T, n_epochs=1, n_microbatches=1, algo="a2c", w_entropy=0.01)
agent = b.OnPolicyAgent(ctx, net, env, cfg, sh.splitmix_states(4, 0, n), host_actions=False)
env.reset_(True)
def it():
    agent.collect(T); agent.update()
ms = timed(it, 10)
print(json.dumps({"bench": "config 3: 32768 Pendulum + A2C-GAE(0.95), Gaussian head, T=32", "ms_per_iteration": ms, "env_steps_per_s": n * T / ms * 1e3,
                  "episodes": env.episode_stats()["episodes"]}), flush=True)
agent.close(); net.close(); env.close()

# ---- config 5: DQN, 1M-transition prioritised replay, batch 4096 --------------------------------
lanes, cap, B = 4096, 256, 4096               # 4096 x 256 = 1,048,576 transitions
env = b.B200VecEnv(ctx, "CartPole", lanes, sh.splitmix_states(7, 0, lanes), auto_reset=True)
tr = b.Trajectory(ctx, 4, cap, lanes=lanes, batch_size=B, sampler_rng=sh.splitmix_states(8, 0, B), prioritized=True, default_priority=1.0)
env.reset_(True); tr.push_env(env, first_state_only=True)
t0 = time.perf_counter()
for _ in range(cap):
    env.act_random_(); tr.push_env(env)
ctx.sync(); fill_s = time.perf_counter() - t0
r = np.random.default_rng(9)
def dense(o, i):
    lim = np.sqrt(6.0 / (i + o)); return [r.uniform(-lim, lim, (o, i)).astype(np.float32).ravel(order="F"), np.zeros(o, np.float32)]
qparams = np.concatenate(dense(128, 4) + dense(128, 128) + dense(2, 128))
qnet = b.Network(ctx, 4, 128, 2, qparams, kind=b.KIND_Q)
learner = b.DQNLearner(ctx, qnet, tr, b.dqn_config(target_update_freq=100))
ms = timed(lambda: learner.update(), 200, warm=10)
st = learner.update(want_stats=True)
ms_sample = timed(lambda: tr.sample(fetch=False), 200)
print(json.dumps({"bench": "config 5: DQN 4-128-128-2, 1,048,576-transition prioritised replay, batch 4096", "ms_per_update": ms, "updates_per_s": 1e3 / ms,
                  "samples_per_s": B / ms * 1e3, "ms_sample_gather_only": ms_sample, "gather_gbs_at_326B": B * 326 / ms_sample / 1e6,
                  "replay_fill_transitions_per_s": lanes * cap / fill_s, "loss": float(st["loss"]), "len": len(tr)}), flush=True)
