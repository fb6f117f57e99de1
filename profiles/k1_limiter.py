"""K1 (env_step_kernel, CartPole{Float32}, given actions) at 16 M envs for an `ncu --set full` capture: which pipe limits it?
    ncu --set full --clock-control none -k regex:env_step_kernel -s 3 -c 1 -o gpurun_out/prof_k1 -f python profiles/k1_limiter.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

b = g.load_package()
ctx = b.Context(0)
n = 1 << 24
env = b.B200VecEnv(ctx, "CartPole", n, b.sharding.splitmix_states(1, 0, n), auto_reset=True)
d_act = ctx.malloc(n * 4)
ctx.h2d(d_act, np.ones(n, np.int32))
for _ in range(6):
    env.act_(d_act)
ctx.sync()
ctx.timer_start()
for _ in range(20):
    env.act_(d_act)
ms = ctx.timer_stop_ms() / 20
print(f"env_step_kernel given actions, {n} envs: {ms:.4f} ms, {n * 57 / ms / 1e6:.0f} GB/s algorithmic (57 B per env-step)")
for _ in range(3):
    env.act_random_()
ctx.sync()
ctx.timer_start()
for _ in range(20):
    env.act_random_()
ms = ctx.timer_stop_ms() / 20
print(f"env_step_kernel fused random policy, {n} envs: {ms:.4f} ms, {n * 117 / ms / 1e6:.0f} GB/s algorithmic (117 B per env-step)")
