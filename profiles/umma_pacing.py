"""Pacing of tcgen05.mma kind::f16 (cycles per instruction, steady state) against the shared-memory layout type, N, M, the operand
source (SS / TS), one vs two accumulators, and spinning vs sleeping bystander threads.  Operand values are zeros (irrelevant).
Build: python reinforcementlearning.jl_b200/build.py --selftest ; run on a B200."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import _selftest as st

b = g.load_package()
ctx = b.Context(0)
lib = st.load()
lib.b200rl_selftest_pacing.restype = C.c_int
lib.b200rl_selftest_pacing.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]


def idesc_f16(M, N):
    return (1 << 4) | ((N >> 3) << 17) | ((M >> 4) << 24)


def once(M, N, layout, lbo, sbo, kadv, n, ts=0, alt=0, sleep=0, elect=0, mn=0):
    p = np.array([idesc_f16(M, N) | (mn << 15) | (mn << 16), layout, lbo, sbo, kadv, n, ts, alt, sleep, elect], np.uint32)
    out = np.zeros(2, np.float32)
    b._lib.check(lib.b200rl_selftest_pacing(ctx.h, p.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    return out


def pace(M, N, layout, lbo, sbo, kadv, **kw):
    once(M, N, layout, lbo, sbo, kadv, 16, **kw)
    lo = once(M, N, layout, lbo, sbo, kadv, 16, **kw)
    hi = once(M, N, layout, lbo, sbo, kadv, 16 + 256, **kw)
    return (hi - lo) / 256.0


LAY = {0: "NONE", 1: "128B_base32B", 2: "128B", 4: "64B", 6: "32B"}
print("cycles per tcgen05.mma kind::f16 K=16 (issue-side, issue+completion), one issuing thread, 256 dependent accumulates")
for elect in (0, 1):
    print("issuing lane chosen by", "elect.sync" if elect else "threadIdx.x == 0 (ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY loop)")
    for layout, lbo, sbo, kadv in ((0, 128, 1024, 256), (2, 16, 1024, 32)):
        for M in (128, 64):
            for N in (16, 64, 128, 144, 256):
                r = pace(M, N, layout, lbo, sbo, kadv, elect=elect)
                print(f"  SS layout={LAY[layout]:13s} M={M:3d} N={N:3d}: issue {r[0]:6.1f}  done {r[1]:6.1f}")
print("elect.sync from here on")
for layout, lbo, sbo, kadv in ((0, 128, 1024, 256), (2, 16, 1024, 32)):
    for N in (64, 128, 256):
        r = pace(128, N, layout, lbo, sbo, kadv, ts=1, elect=1)
        print(f"  TS layout={LAY[layout]:13s} M=128 N={N:3d}: issue {r[0]:6.1f}  done {r[1]:6.1f}")
    for N in (64, 128):
        r = pace(128, N, layout, lbo, sbo, kadv, alt=1, elect=1)
        print(f"  SS layout={LAY[layout]:13s} M=128 N={N:3d} two accumulators: issue {r[0]:6.1f}  done {r[1]:6.1f}")
        r = pace(128, N, layout, lbo, sbo, kadv, sleep=1, elect=1)
        print(f"  SS layout={LAY[layout]:13s} M=128 N={N:3d} bystanders asleep: issue {r[0]:6.1f}  done {r[1]:6.1f}")
    r = pace(128, 128, layout, lbo, sbo, 0, elect=1)
    print(f"  SS layout={LAY[layout]:13s} M=128 N=128 same operand address every time: issue {r[0]:6.1f}  done {r[1]:6.1f}")
print("MN-major A and B (GEMM3 / GEMM4 of K7: LBO = 128 between 8-sample k-blocks, SBO = 2048 between 8-feature blocks), elect.sync")
for N in (16, 64, 144):
    r = pace(128, N, 0, 128, 2048, 256, elect=1, mn=1)
    print(f"  SS MN-major M=128 N={N:3d}: issue {r[0]:6.1f}  done {r[1]:6.1f}")
