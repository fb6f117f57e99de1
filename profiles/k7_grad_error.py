"""Where does the K7 gradient differ from the FP64-accumulating oracle?  Per parameter group (W1, b1, W2, b2, W3, b3 of actor and critic)
relative L2 error of the first update's gradient on a real rollout at BASELINE size.  Usage: python profiles/k7_grad_error.py [pendulum|cartpole]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import oracle_lib as O

pkg = g.load_package()
ctx = pkg.Context(0)
which = sys.argv[1] if len(sys.argv) > 1 else "pendulum"
R = pkg.learners
if which == "pendulum":
    n, T, ns, nout, kind, act, algo, oalgo = 32768, 32, 3, 1, pkg.KIND_GAUSSIAN, 1, "a2c", 1
    env = pkg.B200VecEnv(ctx, "Pendulum", n, O.splitmix_states_fast(n, 31), auto_reset=True)
else:
    n, T, ns, nout, kind, act, algo, oalgo = 65536, 32, 4, 2, pkg.KIND_CATEGORICAL, 0, "ppo", 0
    env = pkg.B200VecEnv(ctx, "CartPole", n, O.splitmix_states_fast(n, 31), auto_reset=True)
desc = O.ac_desc(ns, 64, nout, act, kind == pkg.KIND_GAUSSIAN)
params = O.glorot_params(desc, 8)
params = params + 0.05 * np.random.default_rng(8).standard_normal(params.size).astype(np.float32)
net = pkg.Network(ctx, ns, 64, nout, params, act=act, kind=kind)
cfg = pkg.onpolicy_config(update_freq=T, n_epochs=1, n_microbatches=1, algo=algo, w_entropy=0.01, max_grad_norm=0.0, lr=0.0)
agent = pkg.OnPolicyAgent(ctx, net, env, cfg, O.splitmix_states_fast(n, 32), host_actions=False)
env.reset_(is_force=True)
agent.collect(T)
S, A, LP = agent.rollout(R.ROLL_STATE), agent.rollout(R.ROLL_ACTION), agent.rollout(R.ROLL_LOGP)
stats = agent.update(None, want_stats=True)
ADV, RET = agent.rollout(R.ROLL_ADV), agent.rollout(R.ROLL_RET)
nt = n * T
advf = ADV.ravel(order="F")
mean, inv_std = O.adv_norm(advf)
gref, l = O.ac_loss_grad(oalgo, desc, O.hyper_array(w_entropy=0.01), params, np.asfortranarray(S[:, :, :T]).reshape(ns, nt, order="F"), A.ravel(order="F"),
                         LP.ravel(order="F"), advf, RET.ravel(order="F"), None, mean, inv_std)
got = net.get(R.NET_GRAD).astype(np.float64)       # max_grad_norm = 0: unclipped
gref = gref.astype(np.float64)
H = 64
def groups(n_in, n_out, heads2):
    out, off = [], 0
    for name, sz in (("W1", H * n_in), ("b1", H), ("W2", H * H), ("b2", H)):
        out.append((name, off, off + sz)); off += sz
    hs = n_out * H + n_out
    out.append(("W3b3", off, off + hs)); off += hs
    return out, off
ga, na = groups(ns, 2 if kind == pkg.KIND_GAUSSIAN else nout, kind == pkg.KIND_GAUSSIAN)
gc, nc = groups(ns, 1, False)
print(f"{which}: grad norm gpu {np.linalg.norm(got):.6f} oracle {np.linalg.norm(gref):.6f} rel {abs(np.linalg.norm(got) - np.linalg.norm(gref)) / np.linalg.norm(gref):.2e}; total rel L2 err {np.linalg.norm(got - gref) / np.linalg.norm(gref):.2e}")
for role, gs, base in (("actor", ga, 0), ("critic", gc, na)):
    for name, a, b_ in gs:
        x, y = got[base + a:base + b_], gref[base + a:base + b_]
        print(f"  {role:6s} {name:5s} |g| {np.linalg.norm(y):10.4e}  rel L2 err {np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300):.2e}  signed mean rel {np.mean((x - y)) / max(np.mean(np.abs(y)), 1e-300):+.2e}")
