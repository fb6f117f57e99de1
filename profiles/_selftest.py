"""Loader of the tcgen05 layout probe (reinforcementlearning.jl_b200/build/libb200rl_selftest.so — a diagnostic library built by
`python reinforcementlearning.jl_b200/build.py --selftest`, not part of the product library)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "reinforcementlearning.jl_b200", "build", "libb200rl_selftest.so")


def load():
    lib = C.CDLL(SO, mode=C.RTLD_GLOBAL)
    lib.b200rl_selftest_umma.restype = C.c_int
    lib.b200rl_selftest_umma.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]
    return lib


def run(pkg, ctx, lib, a_img, b_img, desc8, ncols=64):
    """desc8 = [a_lbo, a_sbo, b_lbo, b_sbo, a_kadv, b_kadv, idesc, ksteps | a_from_tmem << 16 | f16 << 17 | a_raw << 18 | repeat << 20]"""
    out = np.zeros((128, ncols), np.float32)
    d8 = np.array(desc8, np.uint32)
    pkg._lib.check(lib.b200rl_selftest_umma(ctx.h, a_img.ctypes.data_as(C.c_void_p), a_img.nbytes, b_img.ctypes.data_as(C.c_void_p), b_img.nbytes,
                                            d8.ctypes.data_as(C.c_void_p), ncols, out.ctypes.data_as(C.c_void_p)))
    return out
