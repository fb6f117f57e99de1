"""Per-phase cycle breakdown of the tensor-core K7 kernel (ac_loss_grad_tc_kernel) as seen by CTA 0 / thread 0.

Build step (here, no GPU):   python profiles/k7_phase_timing.py --build
Run step (on the GPU box):   python profiles/k7_phase_timing.py [tid ...]      (default: the three MMA issuers + one plain warp)
The debug library is the normal one with nn_tc.cu recompiled with -DB200RL_K7_TIMING (clock64 marks); it is
not the product library and is never loaded by the package unless this script swaps it in."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "reinforcementlearning.jl_b200")
DBG = os.path.join(PKG, "build", "libb200rl_dbg.so")
NAMES = ["wait G1 (bar1)", "P3 (D1 -> H2, head partials)", "sync a", "P45 loss + dP2 (compute, tmem_st)", "wait G3 prev (bar3)",
         "P45 image stores (FP, FH)", "sync b", "issue G2", "P0 publish + prefetch", "sync c", "P1 layer 1 + tmem_st", "sync d",
         "issue G1(next) + G3", "wait G2 (bar2)", "P7 (D2 -> dP1, dW1)", "(tile count)", "  P0a: publish (registers -> XT image; waits for the prefetch)", "  P0b: request next records + index",
         "issuer: wait RdyA (P45 done everywhere)", "issuer: issue G2", "issuer: wait RdyB (layer 1 done)", "issuer: issue G1(next) + G3",
         "issuer: wait RdyC (P7 done)", "issuer: issue G4",
         "TAIL (cycles per LAUNCH / tiles): drain + head reductions + partial rows", "tail: threadfence + CTA barrier", "tail: grid barrier A (all CTAs' rows written)",
         "tail: stage partial rows (L2 -> smem)", "tail: column sums + sum of squares", "tail: grid barrier B", "tail: fetch 148 sums, norm, clip scale",
         "tail: Adam + done counter"]


def build():
    sys.path.insert(0, PKG)
    import build as B
    B.build()
    objs = [os.path.join(B.BUILD, s.replace(".cu", ".o")) for s, _ in B.SOURCES + B.OPTIONAL if s != "nn_tc.cu"]
    dbg_o = os.path.join(B.BUILD, "nn_tc_dbg.o")
    nvcc = B._nvcc()
    subprocess.check_call([nvcc, "-ccbin", "/usr/bin/g++"] + B.ARCH + B.COMMON + ["-DB200RL_K7_TIMING", "-c", os.path.join(B.CSRC, "nn_tc.cu"), "-o", dbg_o])
    subprocess.check_call([nvcc, "-ccbin", "/usr/bin/g++"] + B.ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", DBG] + objs + [dbg_o, "-ldl"])
    print(DBG)


def run():
    sys.path.insert(0, ROOT)
    import numpy as np
    import __graft_entry__ as g
    b = g.load_package()
    b._lib._LIB = None                  # swap in the instrumented library before anything is loaded
    b._lib.SO_PATH = DBG
    lib = b._lib.load()
    assert hasattr(lib, "b200rl_debug_k7_phases"), "debug library not loaded"
    ctx = b.Context(0)
    n = 65536
    env = b.B200VecEnv(ctx, "CartPole", n, b.sharding.splitmix_states(1, 0, n), auto_reset=True)
    net = b.Network(ctx, 4, 64, 2, b.sharding.glorot_actor_critic(123, 4, 64, 2))
    agent = b.OnPolicyAgent(ctx, net, env, b.onpolicy_config(update_freq=32), b.sharding.splitmix_states(2, 0, n))
    b.run(agent, env, b.StopAfterNSteps(64), b.DeviceEpisodeStats())     # two full PPO iterations (warm-up)
    out = (C.c_ulonglong * 40)()
    lib.b200rl_debug_k7_phases.argtypes = [C.c_void_p, C.c_int]
    # all MMAs are issued by the dedicated warp 16 (nn_tc.cu); the watched threads are plain workers: 0 / 224 gather (c < 2), 256 does not (c = 2)
    watch = [int(a) for a in sys.argv[1:] if a.isdigit()] or [0, 224, 256, 65536]   # tid | cta << 16 (cta 1 = a critic CTA)
    cols = {}
    for w in watch:
        assert lib.b200rl_debug_k7_watch(w) == 0
        lib.b200rl_debug_k7_phases(out, 1)
        b.run(agent, env, b.StopAfterNSteps(32), b.DeviceEpisodeStats())     # 16 minibatch launches
        lib.b200rl_debug_k7_phases(out, 0)
        v = list(out)
        tiles = max(1, v[15])
        cols[w] = ([x / tiles for x in v], sum(v[:15]) / tiles, tiles)
    print("cycles per 128-sample tile as seen by thread <tid> of CTA 0 (" + ", ".join(f"{w}: {cols[w][2]} tiles" for w in watch) + ")")
    print(f"  {'phase':48s}" + "".join(f"{'tid ' + str(w):>10s}" for w in watch))
    for k in list(range(15)) + [16, 17] + list(range(18, 32)):
        print(f"  {NAMES[k]:48s}" + "".join(f"{cols[w][0][k]:10.0f}" for w in watch))
    print(f"  {'sum of the 15 phases':48s}" + "".join(f"{cols[w][1]:10.0f}" for w in watch))
    print("  (P0 row = loads issued after perm_index; P0a / P0b are its first two parts and are not in the sum)")


if __name__ == "__main__":
    build() if "--build" in sys.argv else run()
