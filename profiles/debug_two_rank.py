"""debug helper: two ranks in one process on cuda:0, scenario chosen by argv[1]"""
import ctypes as C, sys, os, threading, time
if "lazy" not in sys.argv[1]: os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); L = pkg._lib
scen = sys.argv[1]
ctxs = [pkg.Context(0), pkg.Context(0)]
regions = (C.c_void_p * 2)()
for r, ctx in enumerate(ctxs):
    L.check(ctx.lib.b200rl_comm_init(ctx.h, 2, r, None))
    p = C.c_void_p(); L.check(ctx.lib.b200rl_comm_p2p_export(ctx.h, None, C.byref(p))); regions[r] = p
for ctx in ctxs:
    L.check(ctx.lib.b200rl_comm_p2p_attach(ctx.h, regions))
sh = pkg.sharding
n_total, T = 2048, 8
E, M = (1, 1) if "e1" in scen else (2, 2)
cfg = pkg.onpolicy_config(update_freq=T, n_epochs=E, n_microbatches=M, normalize_advantage=("nonorm" not in scen))
params0 = sh.glorot_actor_critic(5, 4, 64, 2)
def make(ctx, lo, hi):
    env = pkg.B200VecEnv(ctx, "CartPole", hi - lo, sh.splitmix_states(11, lo, hi), auto_reset=True)
    net = pkg.Network(ctx, 4, 64, 2, params0.copy())
    agent = pkg.OnPolicyAgent(ctx, net, env, cfg, sh.splitmix_states(12, lo, hi))
    env.reset_(is_force=True)
    return env, net, agent
shards = [make(ctxs[r], *sh.shard_range(n_total, r, 2)) for r in range(2)]
if "notc" in scen:
    L.check(ctxs[0].lib.b200rl_set_tensor_cores(0))
t0 = time.time()
def log(r, msg): print(f"[{time.time()-t0:6.2f}] rank {r}: {msg}", flush=True)
def run(r):
    env, net, agent = shards[r]
    try:
        agent.collect(T); ctxs[r].sync(); log(r, "collected")
        if "seq" in scen and r == 1: time.sleep(1.0)
        agent.update(None, want_stats=False); log(r, "update launched")
        ctxs[r].sync(); log(r, "update done")
    except Exception as e:
        log(r, f"ERROR {e}")
th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
[t.start() for t in th]; [t.join() for t in th]
p = [shards[r][1].get() for r in range(2)]
print("identical:", np.array_equal(p[0], p[1]), "changed:", not np.array_equal(p[0], params0))
