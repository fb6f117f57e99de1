"""The polynomial / range-reduction constants of the Julia `sin` / `cos` restatement, pinned to their published source.

Julia's base/special/trig.jl and rem_pio2.jl are ports of FreeBSD msun (k_sinf.c, k_cosf.c, k_sin.c, k_cos.c, e_rem_pio2.c,
e_rem_pio2f.c) and keep its constants.  msun publishes every constant twice: as a decimal and as the IEEE-754 bit pattern (a hex float
for the float kernels, two 32-bit hex words in the comments of the double kernels).  The decimal literals in oracle/jl_math.hpp and in
the device header csrc/jl_device.cuh must (a) be the same text in both files and (b) denote exactly msun's bit patterns."""
import os
import re
import struct

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = open(os.path.join(ROOT, "oracle", "jl_math.hpp")).read()
DEVICE = open(os.path.join(ROOT, "reinforcementlearning.jl_b200", "csrc", "jl_device.cuh")).read()


def words(hi, lo):
    return struct.unpack(">d", struct.pack(">II", hi, lo))[0]


# (name, decimal literal as it appears in the sources, msun bit pattern)
FLOAT_KERNELS = [   # k_sinf.c / k_cosf.c: "S1 = -0x15555554cbac77.0p-55" ...
    ("S1", "-0.16666666641626524", float.fromhex("-0x15555554cbac77p-55")), ("S2", "0.008333329385889463", float.fromhex("0x111110896efbb2p-59")),
    ("S3", "-0.00019839334836096632", float.fromhex("-0x1a00f9e2cae774p-65")), ("S4", "2.718311493989822e-6", float.fromhex("0x16cd878c3b46a7p-71")),
    ("C0", "-0.499999997251031", float.fromhex("-0x1ffffffd0c5e81p-54")), ("C1", "0.04166662332373906", float.fromhex("0x155553e1053a42p-57")),
    ("C2", "-0.001388676377460993", float.fromhex("-0x16c087e80f1e27p-62")), ("C3", "2.439044879627741e-5", float.fromhex("0x199342e0ee5069p-68")),
]
DOUBLE_KERNELS = [  # k_sin.c / k_cos.c: "S1 = -1.66666666666666324348e-01, /* 0xBFC55555, 0x55555549 */" ...
    ("S1", "-1.66666666666666324348e-01", words(0xBFC55555, 0x55555549)), ("S2", "8.33333333332248946124e-03", words(0x3F811111, 0x1110F8A6)),
    ("S3", "-1.98412698298579493134e-04", words(0xBF2A01A0, 0x19C161D5)), ("S4", "2.75573137070700676789e-06", words(0x3EC71DE3, 0x57B1FE7D)),
    ("S5", "-2.50507602534068634195e-08", words(0xBE5AE5E6, 0x8A2B9CEB)), ("S6", "1.58969099521155010221e-10", words(0x3DE5D93A, 0x5ACFD57C)),
    ("C1", "4.16666666666666019037e-02", words(0x3FA55555, 0x5555554C)), ("C2", "-1.38888888888741095749e-03", words(0xBF56C16C, 0x16C15177)),
    ("C3", "2.48015872894767294178e-05", words(0x3EFA01A0, 0x19CB1590)), ("C4", "-2.75573143513906633035e-07", words(0xBE927E4F, 0x809C52AD)),
    ("C5", "2.08757232129817482790e-09", words(0x3E21EE9E, 0xBDB4B1C4)), ("C6", "-1.13596475577881948265e-11", words(0xBDA8FAE9, 0xBE8838D4)),
]
REM_PIO2 = [        # e_rem_pio2.c (double) and e_rem_pio2f.c (float argument, double arithmetic)
    ("invpio2", "6.36619772367581382433e-01", words(0x3FE45F30, 0x6DC9C883)), ("pio2_1", "1.57079632673412561417e+00", words(0x3FF921FB, 0x54400000)),
    ("pio2_1t", "6.07710050650619224932e-11", words(0x3DD0B461, 0x1A626331)), ("pio2_2", "6.07710050630396597660e-11", words(0x3DD0B461, 0x1A600000)),
    ("pio2_2t", "2.02226624879595063154e-21", words(0x3BA3198A, 0x2E037073)), ("pio2_3", "2.02226624871116645580e-21", words(0x3BA3198A, 0x2E000000)),
    ("pio2_3t", "8.47842766036889956997e-32", words(0x397B839A, 0x252049C1)),
    ("pio2_1 (float)", "1.57079631090164184570e+00", words(0x3FF921FB, 0x50000000)), ("pio2_1t (float)", "1.58932547735281966916e-08", words(0x3E5110B4, 0x611A6263)),
]


@pytest.mark.parametrize("name,literal,bits", FLOAT_KERNELS + DOUBLE_KERNELS + REM_PIO2, ids=lambda v: v if isinstance(v, str) and len(v) < 16 else None)
def test_constant_is_msun_bit_pattern_in_both_sources(name, literal, bits):
    assert float(literal) == bits, f"{name}: {literal} is not msun's constant"
    pat = re.escape(literal.lstrip("-"))
    assert re.search(pat, ORACLE), f"{name} missing from oracle/jl_math.hpp"
    assert re.search(pat, DEVICE), f"{name} missing from csrc/jl_device.cuh"
