#!/usr/bin/env python
"""bench.py — env-steps/s of the full PPO iteration on 65 536 batched CartPole envs
(BASELINE.json metric / configs[1]; configs[3] when --gpus > 1: the same 65 536 envs sharded
by index, one gradient all-reduce per optimiser step).

One "step" = one PPO iteration = rollout of T = 32 env steps for every env (policy inference +
env step + trajectory push fused on the device) + GAE/returns/normalisation + 4 epochs x 4
minibatches of {loss + backward, reduce, clip + Adam}  =  N*T env-steps.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference-shaped CPU arm (oracle port, all host cores)

Prints ONE JSON line (see DESIGN.md "Measurement").  Timing: CUDA events on the launching
stream per step (L2 flushed between steps, outside the timed region), barrier + synchronise
around the loop, max over ranks."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS, T_ROLLOUT, N_EPOCHS, N_MICRO, HIDDEN = 65536, 32, 4, 4, 64
# algorithmic work per sample: MLP forward actor 8 960 + critic 8 832 FLOP; fwd+bwd = 3x
FLOP_FWD = 2 * (4 * 64 + 64 * 64 + 64 * 2) + 2 * (4 * 64 + 64 * 64 + 64 * 1)
FLOP_FWD_BWD = 3 * FLOP_FWD
BYTES_ENV_STEP = 49 + 8          # state 16R+16W, action 4R, t 4R+4W, reward 4W, flag 1W (+ episode-return 4R+4W)
BYTES_GAE = 17 + 8               # r 4, v 4, term 1 read; adv 4, ret 4 write (+ 8 for the normalisation sums in registers: none) -> 25 with returns
METRIC = "env-steps/sec at 65536 CartPole envs (full PPO iteration: rollout T=32 + GAE + 4 epochs x 4 minibatches)"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(pw) if pw else None}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def usable_cores():
    """Host threads this process may really use: scheduler affinity capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return n


def cpu_arm_setup(O, n):
    """Build the oracle-port PPO job on `n` envs and pick the OpenMP thread count that runs it fastest
    (one short calibration iteration per candidate: an over-subscribed or quota-limited box is slower
    with every hardware thread than with fewer).  Returns (step_fn, cores)."""
    env = O.OracleVecEnv(O.KIND_CARTPOLE, n, O.splitmix_states_fast(n, 1))
    env.reset(True)
    desc = O.ac_desc(4, HIDDEN, 2)
    p = O.glorot_params(desc, 123); m = np.zeros_like(p); v = np.zeros_like(p); bt = np.array([0.9, 0.999], np.float32)
    prng = O.splitmix_states_fast(n, 2).copy()
    hyper = O.hyper_array()
    counter = [0]

    def step(T=T_ROLLOUT):
        s, _ = O.ppo_iteration(env, desc, hyper, p, m, v, bt, prng, T, N_EPOCHS, N_MICRO, counter[0])
        counter[0] += 1
        return s

    top = max(1, min(usable_cores(), O.lib().orc_max_threads()))
    cands = sorted({top, max(1, top // 2), max(1, top // 4), min(top, 32), min(top, 16), min(top, 8)}, reverse=True)
    best, best_s = top, None
    for c in cands:
        O.lib().orc_set_threads(c)
        s = min(step(T=8), step(T=8))   # best of two short iterations (8-step rollout + the full update shape) per candidate
        if best_s is None or s < best_s:
            best, best_s = c, s
    O.lib().orc_set_threads(best)
    return step, best


def run_reference(args):
    """The reference's own CPU implementation of the path cannot run here (Julia absent, RLZoo /
    MultiThreadEnv not in the snapshot): this arm times the oracle port — per-env heap objects,
    OpenMP parallel-for over envs (Threads.@threads stand-in), Appendix-B PPO update — on all host
    cores, each step a bounded sample (8192 of the 65 536 envs)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    n = args.cpu_envs
    step, cores = cpu_arm_setup(O, n)
    for w in range(args.warmup):
        step()
    secs = [step() for _ in range(args.steps)]
    total = float(np.sum(secs))
    value = n * T_ROLLOUT * args.steps / total
    sample = f"{n} of {N_ENVS} envs per step (same T, epochs, minibatches)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "PPO 2x64 actor-critic on batched CartPoleEnv{Float32}, T=32, 4 epochs x 4 minibatches", "n_envs": n,
                   "parallelism": f"cpu{cores}", "note": "oracle port of the reference path; bounded sample"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_own(args):
    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    import __graft_entry__ as g
    pkg = g.load_package()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(local_rank)
    peer_exchange = False
    if world > 1:
        import ctypes as C
        import torch
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (C.c_char * 128)()
            pkg._lib.check(ctx.lib.b200rl_comm_unique_id(raw))
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        raw = bytes(idbuf.cpu().numpy().tobytes())
        pkg._lib.check(ctx.lib.b200rl_comm_init(ctx.h, world, rank, raw))

        def all_gather_bytes(b):
            t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [bytes(o.cpu().numpy().tobytes()) for o in out]
        peer_exchange = False
        if os.environ.get("B200RL_P2P", "1") != "0":
            from b200rl import sharding as _sh
            peer_exchange = _sh.attach_peer_exchange(ctx, rank, world, all_gather_bytes)

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize()

    n_total = args.envs
    n = n_total // world
    T = T_ROLLOUT
    from b200rl import sharding
    lo, hi = sharding.shard_range(n_total, rank, world)
    states = lambda seed: sharding.splitmix_states(seed, lo, hi)      # streams keyed by GLOBAL env index
    glorot = lambda seed: sharding.glorot_actor_critic(seed, 4, HIDDEN, 2)

    cfg = pkg.onpolicy_config(update_freq=T, n_epochs=N_EPOCHS, n_microbatches=N_MICRO)

    def make(host_actions):
        env = pkg.B200VecEnv(ctx, "CartPole", n, states(0x9E3779B97F4A7C15), auto_reset=True)
        net = pkg.Network(ctx, 4, HIDDEN, 2, glorot(123))
        agent = pkg.OnPolicyAgent(ctx, net, env, cfg, states(0x1234567), host_actions=host_actions)
        env.reset_(is_force=True)
        return env, net, agent

    env, net, agent = make(False)
    for _ in range(args.warmup):
        agent.collect(T)
        agent.update()
    clocks = ClockSampler(local_rank)
    launches0 = ctx.launch_count()
    barrier()
    clocks.start()
    total_ms = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.timer_start()
        agent.collect(T)
        agent.update()
        total_ms += ctx.timer_stop_ms()
    barrier()
    clk = clocks.stop()
    launches = ctx.launch_count() - launches0
    if dist is not None:
        import torch
        tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = n_total * T * args.steps / (total_ms / 1000.0)
    stats = env.episode_stats()

    # ---- e2e: the reference-facing stage protocol with HOST action buffers every step ---------
    e2e = None
    if not args.no_e2e:
        env2, net2, agent2 = make(True)
        agent2.fetch_stats = True
        k_e2e = max(1, min(args.steps, 5))
        pkg.run(agent2, env2, pkg.StopAfterNSteps(T * max(1, min(args.warmup, 2))), pkg.DeviceEpisodeStats())
        barrier()
        t0 = time.perf_counter()
        pkg.run(agent2, env2, pkg.StopAfterNSteps(T * k_e2e), pkg.DeviceEpisodeStats())
        ctx.sync()
        sec = time.perf_counter() - t0
        if dist is not None:
            import torch
            tt = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec = float(tt.item())
        e2e = {"value": n_total * T * k_e2e / sec, "unit": "env-steps/s", "h2d_bytes_per_step": int(T * n * 4),
               "d2h_bytes_per_step": int(T * n * 4 + N_EPOCHS * N_MICRO * 6 * 4 + 32), "steps": k_e2e,
               "note": "run(agent, env, StopAfterNSteps, hook): plan! returns host actions, act! takes host actions, per-update loss read-back"}
        agent2.close(); net2.close(); env2.close()

    # ---- roofline of the dominant kernel + per-kernel times (CUDA events, back-to-back launches) ----
    hbm_peak, tf_peak, peak_kind = measured_peaks()
    B_local = n * T // N_MICRO
    k_loss = agent.time_kernel(0, 10)
    k_act = agent.time_kernel(1, 20)
    tc_on = os.environ.get("B200RL_TC", "1") != "0"
    k_loss_ffma = k_act_ffma = None
    if tc_on:   # the FP32 CUDA-core variants of the same kernels, for the record
        pkg._lib.check(ctx.lib.b200rl_set_tensor_cores(0))
        k_loss_ffma = agent.time_kernel(0, 10)
        k_act_ffma = agent.time_kernel(1, 20)
        pkg._lib.check(ctx.lib.b200rl_set_tensor_cores(1))
    k_gae = agent.time_kernel(3, 20)
    k_adam = agent.time_kernel(4, 20)
    k_env = agent.time_kernel(2, 20)
    ach_tf = B_local * FLOP_FWD_BWD / (k_loss * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("ac_loss_grad_tc_kernel" if tc_on else "ac_loss_grad_kernel", tj.get("ac_loss_grad_kernel"))
        except Exception:
            traffic = None
    share = (N_EPOCHS * N_MICRO * k_loss) / (total_ms / args.steps)
    kname = ("ac_loss_grad_tc_kernel (PPO loss + backward, one minibatch; 64x64 GEMMs on tcgen05 as 3xTF32, FP32 accumulate in TMEM)"
             if tc_on else "ac_loss_grad_kernel<64> (PPO loss + backward, one minibatch; FP32 FFMA)")
    roofline = {"kernel": kname, "bound": "tensor", "achieved": ach_tf, "peak": tf_peak,
                "unit": "TFLOP/s", "frac": ach_tf / tf_peak, "traffic": traffic, "peak_kind": f"bf16 dense GEMM burst, {peak_kind}",
                "note": "achieved = algorithmic FP32 FLOPs (53,376 per sample) / event time; the 1e-5 parity bar needs 3xTF32 (3 tensor-core products per "
                        "algorithmic product, K = 8 per instruction) so the executed tensor work is 3x the algorithmic; the same figure as a fraction of the "
                        "FP32 CUDA-core peak (~72 TFLOP/s @1.9 GHz) = %.3f" % (ach_tf / 72.0),
                "ms_per_launch": k_loss, "ms_per_launch_fp32_ffma_variant": k_loss_ffma, "share_of_step": share,
                "other_kernels": {
                    "policy_act_ms": k_act, "policy_act_ms_fp32_ffma_variant": k_act_ffma, "policy_act_tflops": n * FLOP_FWD / (k_act * 1e-3) / 1e12,
                    "env_step_ms": k_env, "env_step_gbs": n * BYTES_ENV_STEP / (k_env * 1e-3) / 1e9, "env_step_frac_hbm": n * BYTES_ENV_STEP / (k_env * 1e-3) / 1e9 / hbm_peak,
                    "gae_ms": k_gae, "gae_gbs": n * T * BYTES_GAE / (k_gae * 1e-3) / 1e9, "gae_frac_hbm": n * T * BYTES_GAE / (k_gae * 1e-3) / 1e9 / hbm_peak,
                    "reduce_clip_adam_ms": k_adam}}

    # ---- CPU baseline (rank 0, N = 1 only): oracle port, bounded sample ------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        nc = args.cpu_envs
        step, cores = cpu_arm_setup(O, nc)
        s0 = step()
        reps = int(max(2, min(20, 12.0 / max(s0, 1e-3))))
        secs = [step() for _ in range(reps)]
        cpu = {"value": nc * T * reps / float(np.sum(secs)), "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": f"{reps} PPO iterations on {nc} of {N_ENVS} envs (same T / epochs / minibatches), OpenMP over per-env heap objects, thread count picked by calibration"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: 65536 batched CartPoleEnv{Float32} + PPO (2x64 MLP actor-critic, fp32), T=32, 4 epochs x 4 minibatches, advantage normalisation",
                       "n_envs": n_total, "envs_per_gpu": n, "global_batch": n_total * T // N_MICRO, "parallelism": f"dp{world}",
                       "grad_allreduce": ("fused NVLink peer exchange inside the reduce+clip+Adam kernel" if peer_exchange else "NCCL") if world > 1 else "none",
                       "l2": "flushed (256 MB write) between timed steps, outside the timed region",
                       "episodes_finished_rank0": stats["episodes"]},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    agent.close(); net.close(); env.close()
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--envs", type=int, default=N_ENVS)
    ap.add_argument("--cpu-envs", type=int, default=8192)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
