"""Host mirror of the learner side: network handle (FluxApproximator + TargetNetwork,
RLCore/src/policies/learners/flux_approximator.jl:11-46, target_network.jl:27-88), the
on-policy agent (Agent + PPOPolicy/A2C, RLCore/src/policies/agent/agent_base.jl:18-66 and the
absent RLZoo learners), the device trajectory (CircularArraySARTSTraces + samplers) and the
DQN learner.  Thin ctypes calls only."""
import ctypes as C

import numpy as np

from . import _lib as L
from .core import AbstractPolicy, FusedAction, PostActStage, PreActStage, PreEpisodeStage, PreExperimentStage

ACT_RELU, ACT_TANH = 0, 1
KIND_CATEGORICAL, KIND_GAUSSIAN, KIND_Q = 0, 1, 2
NET_PARAMS, NET_GRAD, NET_M, NET_V, NET_BETA_T, NET_TARGET = range(6)


def onpolicy_config(gamma=0.99, lambda_=0.95, clip_range=0.1, max_grad_norm=0.5, w_actor=1.0, w_critic=0.5, w_entropy=0.001,
                    lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, min_sigma=0.0, max_sigma=float("inf"), normalize_advantage=True,
                    n_epochs=4, n_microbatches=4, update_freq=32, algo="ppo"):
    """Defaults = the in-tree PPO example (docs/homepage/blog/a_practical_introduction_to_RL.jl/index.html:15238-15286)."""
    return L.OnPolicyConfig(gamma, lambda_, clip_range, max_grad_norm, w_actor, w_critic, w_entropy, lr, beta1, beta2, eps, min_sigma,
                            max_sigma, int(normalize_advantage), n_epochs, n_microbatches, update_freq, {"ppo": 0, "a2c": 1}[algo])


def dqn_config(gamma=0.99, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, max_grad_norm=0.0, rho=0.0, per_alpha=0.6, per_beta=0.4,
               per_eps=1e-6, huber=True, double_dqn=False, target_update_freq=100):
    return L.DQNConfig(gamma, lr, beta1, beta2, eps, max_grad_norm, rho, per_alpha, per_beta, per_eps, int(huber), int(double_dqn),
                       target_update_freq)


class Network:
    """b200rl_net: parameters + Adam state (+ target copy) on the device."""

    def __init__(self, ctx, n_in, hidden, n_out, params, act=ACT_RELU, kind=KIND_CATEGORICAL):
        self.ctx, self.lib = ctx, ctx.lib
        self.desc = L.NetDesc(n_in, hidden, act, n_out, kind)
        n = C.c_int64()
        L.check(self.lib.b200rl_net_nparams(C.byref(self.desc), C.byref(n)))
        self.nparams = n.value
        params = np.ascontiguousarray(params, dtype=np.float32)
        if params.size != self.nparams:
            raise ValueError(f"expected {self.nparams} parameters, got {params.size}")
        h = C.c_void_p()
        L.check(self.lib.b200rl_net_create(ctx.h, C.byref(self.desc), L.ptr(params), C.byref(h)))
        self.h = h
        self.kind, self.n_in, self.n_out = kind, n_in, n_out

    @staticmethod
    def count_params(ctx, n_in, hidden, n_out, act=ACT_RELU, kind=KIND_CATEGORICAL):
        d = L.NetDesc(n_in, hidden, act, n_out, kind)
        n = C.c_int64()
        L.check(ctx.lib.b200rl_net_nparams(C.byref(d), C.byref(n)))
        return n.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200rl_net_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure_optimizer(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, max_grad_norm=0.5):
        L.check(self.lib.b200rl_net_configure_optimizer(self.h, lr, beta1, beta2, eps, max_grad_norm))

    def get(self, which=NET_PARAMS):
        out = np.empty(2 if which == NET_BETA_T else self.nparams, np.float32)
        L.check(self.lib.b200rl_net_get(self.h, which, L.ptr(out), out.size))
        return out

    def set(self, which, arr):
        arr = np.ascontiguousarray(arr, np.float32)
        L.check(self.lib.b200rl_net_set(self.h, which, L.ptr(arr), arr.size))

    def device_ptr(self, which=NET_PARAMS):
        p = C.c_void_p()
        L.check(self.lib.b200rl_net_ptr(self.h, which, C.byref(p)))
        return p.value

    def step_count(self):
        n = C.c_int64()
        L.check(self.lib.b200rl_net_get_step(self.h, C.byref(n)))
        return n.value

    def set_step_count(self, n):
        L.check(self.lib.b200rl_net_set_step(self.h, int(n)))

    def target_sync(self, rho=0.0):
        L.check(self.lib.b200rl_net_target_sync(self.h, rho))

    def act(self, obs, rng_dev):
        """obs (n_in, N) host array; rng_dev: device pointer to (4, N) policy streams."""
        obs = np.asfortranarray(obs, np.float32)
        n = obs.shape[1]
        action = np.empty(n, np.float32 if self.kind == KIND_GAUSSIAN else np.int32)
        logp = np.empty(n, np.float32); value = np.empty(n, np.float32)
        nh = 2 if self.kind == KIND_GAUSSIAN else self.n_out
        heads = np.empty((nh, n), np.float32, order="F")
        L.check(self.lib.b200rl_net_act(self.h, L.ptr(obs), n, C.c_void_p(rng_dev), L.ptr(action), L.ptr(logp), L.ptr(value), L.ptr(heads), 0))
        return dict(action=action, logp=logp, value=value, heads=heads)

    def values(self, obs, use_target=False):
        obs = np.asfortranarray(obs, np.float32)
        n = obs.shape[1]
        out = np.empty((self.n_out, n), np.float32, order="F") if self.kind == KIND_Q else np.empty(n, np.float32)
        L.check(self.lib.b200rl_net_values(self.h, L.ptr(obs), n, L.ptr(out), int(use_target), 0))
        return out

    def ac_step(self, cfg, states, actions, logp_old, adv, ret, idx=None, adv_mean=0.0, adv_inv_std=1.0, apply_update=True):
        states = np.asfortranarray(states, np.float32)
        total = states.shape[1]
        actions = np.ascontiguousarray(actions)
        assert actions.dtype in (np.int32, np.float32)
        idx = None if idx is None else np.ascontiguousarray(idx, np.int32)
        losses = np.zeros(6, np.float32)
        L.check(self.lib.b200rl_net_ac_step(
            self.h, C.byref(cfg), L.ptr(states), L.ptr(actions), L.ptr(None if logp_old is None else np.ascontiguousarray(logp_old, np.float32)),
            L.ptr(np.ascontiguousarray(adv, np.float32)), L.ptr(np.ascontiguousarray(ret, np.float32)), total, L.ptr(idx),
            0 if idx is None else idx.size, adv_mean, adv_inv_std, int(apply_update), L.ptr(losses)))
        return dict(actor_loss=losses[0], critic_loss=losses[1], entropy=losses[2], loss=losses[3], grad_norm=losses[4])


ROLL_STATE, ROLL_ACTION, ROLL_LOGP, ROLL_REWARD, ROLL_TERMINAL, ROLL_VALUE, ROLL_ADV, ROLL_RET, ROLL_RNG, ROLL_NORM = range(10)


class OnPolicyAgent(AbstractPolicy):
    """Agent(policy = PPOPolicy | A2CPolicy, trajectory = PPOTrajectory) on the device.

    Two ways to drive it: the reference's stage protocol (``plan`` returns host actions, the run
    loop calls ``env.act_``, ``push`` / ``optimise`` follow — host buffers every step), or the fused
    path (``collect(n)`` / ``update()``), where actions never leave the device."""

    def __init__(self, ctx, net, env, cfg, policy_rng, host_actions=True):
        self.ctx, self.lib, self.net, self.env, self.cfg = ctx, ctx.lib, net, env, cfg
        self.n, self.T = env.n, cfg.update_freq
        self.host_actions = host_actions
        self.fusable = not host_actions     # run() may hand whole stretches of env steps to collect()
        policy_rng = np.ascontiguousarray(policy_rng, np.uint64).reshape(self.n, 4)
        h = C.c_void_p()
        L.check(self.lib.b200rl_onpolicy_create(ctx.h, net.h, env.h, C.byref(cfg), L.ptr(policy_rng), C.byref(h)))
        self.h = h
        self.continuous = env.continuous
        # pinned host buffer for the per-step action round trip of the stage protocol
        self._act_buf, self._act_buf_addr = ctx.host_alloc((self.n,), np.float32 if self.continuous else np.int32)
        env.pinned_action_addr = self._act_buf.ctypes.data
        self.last_stats = None
        self.n_updates = 0
        self.fetch_stats = False   # read the per-minibatch losses back after every update (a sync)
        self._t = 0                # host mirror of the rollout fill level (no device query per step)

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200rl_onpolicy_destroy(self.h)
            self.h = None
            self._act_buf = None
            self.ctx.host_free(self._act_buf_addr)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def time_kernel(self, which, reps=20):
        ms = C.c_float()
        L.check(self.lib.b200rl_onpolicy_time_kernel(self.h, which, reps, C.byref(ms)))
        return ms.value

    # ---- stage protocol (run.jl:52-68) -------------------------------------------------------
    def plan(self, env):
        if self.host_actions:
            L.check(self.lib.b200rl_onpolicy_plan(self.h, L.ptr(self._act_buf)))
            if self.continuous:
                lo, hi = env.action_space()               # the env asserts a in -2.0..2.0 (Pendulum) | -1.0..1.0
                return np.clip(self._act_buf, lo, hi)
            return self._act_buf
        L.check(self.lib.b200rl_onpolicy_plan(self.h, None))
        return FusedAction("policy")

    def act_fused(self, env):
        L.check(self.lib.b200rl_onpolicy_act(self.h))

    def push(self, stage, env, action=None):
        if stage == PostActStage:
            L.check(self.lib.b200rl_onpolicy_push(self.h))

    def optimise(self, stage):
        if stage == PostActStage:
            self._t += 1
            if self._t == self.T:
                self._t = 0
                self.update(want_stats=self.fetch_stats)

    # ---- fused path --------------------------------------------------------------------------
    def collect(self, n_steps):
        L.check(self.lib.b200rl_onpolicy_collect(self.h, n_steps))
        self._t = (self._t + n_steps)

    def fill(self):
        t, T = C.c_int(), C.c_int()
        L.check(self.lib.b200rl_onpolicy_fill(self.h, C.byref(t), C.byref(T)))
        return t.value, T.value

    def update(self, perm=None, want_stats=False):
        rows = self.cfg.n_epochs * self.cfg.n_microbatches
        stats = np.zeros((rows, 6), np.float32) if want_stats else None
        if perm is not None:
            perm = np.ascontiguousarray(perm, np.int32)
            assert perm.shape == (self.cfg.n_epochs, self.n * self.T)
        L.check(self.lib.b200rl_onpolicy_update(self.h, L.ptr(perm), L.ptr(stats)))
        self._t = 0
        self.n_updates += 1
        self.last_stats = stats
        return stats

    def iterate(self, n_iters=1, want_stats=False):
        """n_iters x {collect(T); update()} as one CUDA-graph launch per iteration (b200rl_onpolicy_iterate)."""
        rows = self.cfg.n_epochs * self.cfg.n_microbatches
        stats = np.zeros((rows, 6), np.float32) if want_stats else None
        L.check(self.lib.b200rl_onpolicy_iterate(self.h, n_iters, L.ptr(stats)))
        self._t = 0
        self.n_updates += n_iters
        self.last_stats = stats
        return stats

    def graph_active(self):
        v = C.c_int()
        L.check(self.lib.b200rl_onpolicy_graph_active(self.h, C.byref(v)))
        return bool(v.value)

    def rollout(self, field):
        n, T, ns = self.n, self.T, self.net.n_in
        spec = {
            ROLL_STATE: ((ns, n, T + 1), np.float32), ROLL_ACTION: ((n, T), np.float32 if self.continuous else np.int32),
            ROLL_LOGP: ((n, T), np.float32), ROLL_REWARD: ((n, T), np.float32), ROLL_TERMINAL: ((n, T), np.uint8),
            ROLL_VALUE: ((n, T + 1), np.float32), ROLL_ADV: ((n, T), np.float32), ROLL_RET: ((n, T), np.float32),
            ROLL_RNG: ((4, n), np.uint64), ROLL_NORM: ((2,), np.float32),
        }[field]
        out = np.empty(spec[0], dtype=spec[1], order="F")
        L.check(self.lib.b200rl_onpolicy_get(self.h, field, L.ptr(out), out.nbytes))
        return out


BATCH_STATE, BATCH_ACTION, BATCH_REWARD, BATCH_TERMINAL, BATCH_NEXT_STATE, BATCH_KEY, BATCH_PRIORITY, BATCH_WEIGHT, BATCH_RNG = range(9)


class Trajectory:
    """Trajectory(container = CircularArraySARTSTraces(capacity) [+ CircularPrioritizedTraces],
    sampler = BatchSampler(batch_size)) resident on the device."""

    def __init__(self, ctx, ns, capacity, lanes=1, batch_size=0, sampler_rng=None, prioritized=False, default_priority=1.0):
        self.ctx, self.lib = ctx, ctx.lib
        self.ns, self.lanes, self.capacity, self.batch_size, self.prioritized = ns, lanes, capacity, batch_size, prioritized
        if batch_size:
            sampler_rng = np.ascontiguousarray(sampler_rng, np.uint64).reshape(batch_size, 4)
        h = C.c_void_p()
        L.check(self.lib.b200rl_traj_create(ctx.h, ns, lanes, capacity, int(prioritized), default_priority, L.ptr(sampler_rng), batch_size, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200rl_traj_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        n = C.c_int64()
        L.check(self.lib.b200rl_traj_length(self.h, C.byref(n)))
        return n.value

    def push_state(self, obs):
        obs = np.asfortranarray(obs, np.float32)
        L.check(self.lib.b200rl_traj_push_state(self.h, L.ptr(obs), 0))

    def push_episode_start(self, obs, pending_only=False, on_device=False):
        """push!(trajectory, (state = s0,)): every lane, or (pending_only) the lanes whose last transition was terminal."""
        if not on_device:
            obs = np.asfortranarray(obs, np.float32)
        L.check(self.lib.b200rl_traj_push_episode_start(self.h, L.ptr(obs), int(on_device), int(pending_only)))

    def lane_lengths(self):
        out = np.empty(self.lanes, np.int64)
        L.check(self.lib.b200rl_traj_lane_lengths(self.h, L.ptr(out)))
        return out

    def n_sampleable(self):
        n = C.c_int64()
        L.check(self.lib.b200rl_traj_n_sampleable(self.h, C.byref(n)))
        return n.value

    # ---- checkpoint of the ring (fields of b200rl_traj_get) ----------------------------------
    FIELDS = {"state": (0, np.float32), "action": (1, np.int32), "reward": (2, np.float32), "flag": (3, np.uint8), "head": (4, np.int32),
              "count": (5, np.int32), "pending": (6, np.uint8), "n_sampleable": (7, np.int64), "tree": (8, np.float32), "sampler_rng": (9, np.uint64)}

    def export_state(self):
        out = {}
        for name, (f, dt) in self.FIELDS.items():
            if (name == "tree" and not self.prioritized) or (name == "sampler_rng" and not self.batch_size):
                continue
            nb = C.c_size_t()
            L.check(self.lib.b200rl_traj_field_bytes(self.h, f, C.byref(nb)))
            a = np.empty(nb.value // np.dtype(dt).itemsize, dt)
            L.check(self.lib.b200rl_traj_get(self.h, f, L.ptr(a), a.nbytes))
            out[name] = a
        if hasattr(self, "controller"):
            c = self.controller
            out["controller"] = np.array([c.ratio, c.threshold, c.n_inserted, c.n_sampled], np.float64)
        return out

    def import_state(self, st):
        for name, (f, dt) in self.FIELDS.items():
            if name in st:
                a = np.ascontiguousarray(st[name], dt)
                L.check(self.lib.b200rl_traj_set(self.h, f, L.ptr(a), a.nbytes))
        if "controller" in st and hasattr(self, "controller"):
            c = self.controller
            c.ratio, c.threshold, c.n_inserted, c.n_sampled = float(st["controller"][0]), int(st["controller"][1]), int(st["controller"][2]), int(st["controller"][3])

    def push(self, action, reward, terminal, next_obs):
        L.check(self.lib.b200rl_traj_push(self.h, L.ptr(np.ascontiguousarray(action, np.int32)), L.ptr(np.ascontiguousarray(reward, np.float32)),
                                          L.ptr(np.ascontiguousarray(terminal, np.uint8)), L.ptr(np.asfortranarray(next_obs, np.float32)), 0))

    def push_env(self, env, first_state_only=False):
        L.check(self.lib.b200rl_traj_push_env(self.h, env.h, int(first_state_only)))

    def sample(self, beta=0.4, fetch=True):
        L.check(self.lib.b200rl_traj_sample(self.h, beta))
        return self.batch() if fetch else None

    def batch(self):
        B, ns = self.batch_size, self.ns
        spec = {"state": (BATCH_STATE, (ns, B), np.float32), "action": (BATCH_ACTION, (B,), np.int32), "reward": (BATCH_REWARD, (B,), np.float32),
                "terminal": (BATCH_TERMINAL, (B,), np.uint8), "next_state": (BATCH_NEXT_STATE, (ns, B), np.float32),
                "key": (BATCH_KEY, (B,), np.int64), "priority": (BATCH_PRIORITY, (B,), np.float32), "weight": (BATCH_WEIGHT, (B,), np.float32)}
        out = {}
        for k, (f, shape, dt) in spec.items():
            a = np.empty(shape, dt, order="F")
            L.check(self.lib.b200rl_traj_batch_get(self.h, f, L.ptr(a), a.nbytes))
            out[k] = a
        return out

    def sampler_rng(self):
        a = np.empty((self.batch_size, 4), np.uint64)
        L.check(self.lib.b200rl_traj_batch_get(self.h, BATCH_RNG, L.ptr(a), a.nbytes))
        return a

    def update_priority(self, prio):
        prio = np.ascontiguousarray(prio, np.float32)
        L.check(self.lib.b200rl_traj_update_priority(self.h, L.ptr(prio), 0))

    def total_priority(self):
        v = C.c_float()
        L.check(self.lib.b200rl_traj_total_priority(self.h, C.byref(v)))
        return v.value


class DQNLearner:
    """DQNLearner / PrioritizedDQNLearner update on a device Trajectory."""

    def __init__(self, ctx, net, traj, cfg):
        self.ctx, self.lib, self.net, self.traj, self.cfg = ctx, ctx.lib, net, traj, cfg

    def update(self, want_stats=False):
        stats = np.zeros(4, np.float32) if want_stats else None
        L.check(self.lib.b200rl_dqn_update(self.net.h, self.traj.h, C.byref(self.cfg), L.ptr(stats)))
        return None if stats is None else dict(loss=stats[0], grad_norm=stats[1], mean_abs_td=stats[2], n_updates=int(stats[3]))

    def last_td(self):
        out = np.empty(self.traj.batch_size, np.float32)
        L.check(self.lib.b200rl_dqn_last_td(self.net.h, self.traj.h, L.ptr(out), out.size))
        return out


class InsertSampleRatioController:
    """InsertSampleRatioController(ratio, threshold) (ReinforcementLearningTrajectories 0.4, external; described in
    docs/src/How_to_implement_a_new_algorithm.md:108): counts insertions and sampled batches; a batch may be sampled once
    ``threshold`` insertions happened and while ``n_sampled <= (n_inserted - threshold) * ratio``.  One insertion = one
    push of a frame (all lanes), the batched counterpart of one ``push!``."""

    def __init__(self, ratio=1.0, threshold=1, n_inserted=0, n_sampled=0):
        self.ratio, self.threshold, self.n_inserted, self.n_sampled = float(ratio), int(threshold), int(n_inserted), int(n_sampled)

    def on_insert(self, n=1):
        self.n_inserted += n

    def on_sample(self):
        if self.n_inserted >= self.threshold and self.n_sampled <= (self.n_inserted - self.threshold) * self.ratio:
            self.n_sampled += 1
            return True
        return False


class QBasedPolicy(AbstractPolicy):
    """QBasedPolicy(learner = DQNLearner(...), explorer = EpsilonGreedyExplorer(...)) (q_based_policy.jl:13-49) on a batched
    env: ``plan`` = BatchExplorer over Q(state(env), .) — forward pass, schedule, draws and arg-max in one device call.

    ``explorer_rng``: (N, 4) uint64 raw Xoshiro states, one explorer stream per env."""

    def __init__(self, ctx, learner, explorer, explorer_rng, n_envs):
        self.ctx, self.lib, self.learner, self.explorer, self.n = ctx, ctx.lib, learner, explorer, int(n_envs)
        rng = np.ascontiguousarray(explorer_rng, np.uint64).reshape(self.n, 4)
        self._d_rng = ctx.malloc(rng.nbytes)
        ctx.h2d(self._d_rng, rng)
        self._d_action = ctx.malloc(self.n * 4)

    def close(self):
        for name in ("_d_rng", "_d_action"):
            p = getattr(self, name, None)
            if p:
                self.ctx.free(p)
                setattr(self, name, None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def plan_device(self, env):
        """plan!(policy, env) leaving the actions on the device; returns the device pointer of the (N,) int32 actions."""
        net, ex = self.learner.net, self.explorer
        obs = C.c_void_p(env.device_ptr(L.FIELD_OBS))
        if hasattr(ex, "as_struct"):
            st = ex.as_struct()
            L.check(self.lib.b200rl_net_q_explore(net.h, obs, self.n, C.c_void_p(self._d_rng), C.byref(st), C.c_void_p(self._d_action)))
            ex.advance(self.n)
        else:   # GreedyExplorer: findmax, no draw
            L.check(self.lib.b200rl_net_q_act(net.h, obs, self.n, None, 0.0, C.c_void_p(self._d_action)))
        return self._d_action

    def plan(self, env):
        self.plan_device(env)
        return FusedAction("policy")

    def act_fused(self, env):
        env.act_(int(self._d_action))

    def explorer_rng(self):
        out = np.empty((self.n, 4), np.uint64)
        return self.ctx.d2h(out, self._d_rng)

    def set_explorer_rng(self, rng):
        self.ctx.h2d(self._d_rng, np.ascontiguousarray(rng, np.uint64).reshape(self.n, 4))

    def optimise(self, stage, trajectory=None):
        """optimise!(policy, stage, trajectory) = optimise!(policy.learner, stage, trajectory) (q_based_policy.jl:48-49);
        the DQN learner trains at the PostActStage."""
        if stage == PostActStage and trajectory is not None:
            while trajectory.controller.on_sample():
                self.learner.update()


class Agent(AbstractPolicy):
    """Agent(policy, trajectory) (agent_base.jl:18-66) for a device-resident replay trajectory: pushes the env's
    transition frames (state / action / reward / terminal never visit the host) and lets the policy's learner train
    whenever the trajectory's controller allows a batch.

    Episode starts (the PreEpisodeStage push of agent_base.jl:45-47): run() announces every forced reset with a PreEpisodeStage
    push -> every lane gets an episode-start frame (so re-entering run() on a filled trajectory is fine: the entry straddling the
    reset exists and is not sampleable, exactly EpisodesBuffer's bookkeeping).  Episodes that end inside the loop: with the env's
    in-kernel auto-reset the trajectory writes the episode-start frame itself (the terminal step's observation already is the new
    episode's first state); with soft resets (auto_reset = False) the lanes whose last transition was terminal get the post-reset
    observation at the next PreActStage."""

    def __init__(self, policy, trajectory, host_actions=False):
        self.policy, self.trajectory, self.host_actions = policy, trajectory, host_actions
        if not hasattr(trajectory, "controller"):
            trajectory.controller = InsertSampleRatioController()
        self._host_act = None

    def push(self, stage, env, action=None):
        if stage == PreEpisodeStage:
            self.trajectory.push_env(env, first_state_only=True)         # push!(trajectory, (state = state(env),)) for every lane
        elif stage == PreActStage:
            if not getattr(env, "auto_reset", True):
                self.trajectory.push_env(env, first_state_only=2)        # lanes that were soft-reset since their terminal transition
        elif stage == PostActStage:
            self.trajectory.push_env(env)                                # (state = s', action, reward, terminal)
            self.trajectory.controller.on_insert(1)

    def plan(self, env):
        if self.host_actions:   # the reference's stage protocol: the action visits the host, the run loop calls act!(env, a)
            d = self.policy.plan_device(env)
            if self._host_act is None:
                self._host_act = np.empty(self.policy.n, np.int32)
            return self.policy.ctx.d2h(self._host_act, d)
        return self.policy.plan(env)

    def act_fused(self, env):
        self.policy.act_fused(env)

    def optimise(self, stage):
        self.policy.optimise(stage, self.trajectory)
