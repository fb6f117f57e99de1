"""Probe: tcgen05.mma kind::f16 with MN-major SWIZZLE_NONE operands in shared memory (A: M contiguous, B: N contiguous).
Hypothesis (cute canonical layout, units of 16 B): element (mn, k) at (mn/8)*SBO + (mn%8)*2 + (k%8)*16 + (k/8)*LBO — a core matrix is
8 k-rows x 16 B (8 consecutive mn elements).  Both assignments of the two strides to the descriptor's LBO / SBO fields are tried.
K7 wants this form for GEMM3 / GEMM4 (reduction over samples): thread = sample then writes 8 consecutive FEATURES as one 16-byte
vector instead of eight 2-byte transposed stores."""
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
rng = np.random.default_rng(0)
F16 = 1 << 17


def idesc_f16(M, N, a_mn, b_mn):
    return (1 << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def image_mn(mat, mn_stride, k_stride, total_bytes):
    """mat[mn, k] -> byte (mn/8)*mn_stride + (mn%8)*2 + (k%8)*16 + (k/8)*k_stride"""
    img = np.zeros(total_bytes // 2, np.float16)
    R, K = mat.shape
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    img[((r // 8) * mn_stride + (r % 8) * 2 + (k % 8) * 16 + (k // 8) * k_stride) // 2] = mat
    return img.view(np.float32)


K = 64
A = rng.integers(-4, 5, (128, K)).astype(np.float32)
for N in (64, 144, 16):
    Bm = rng.integers(-4, 5, (N, K)).astype(np.float32)
    ref = A @ Bm.T
    KS, MS = 128, 128 * (K // 8) + 0      # k-block stride 128 B (contiguous core matrices along k), mn-block stride = K/8 core matrices
    a_img = image_mn(A, MS, KS, 16 * MS)
    b_img = image_mn(Bm, MS, KS, (N // 8) * MS)
    ncols = 64 if N >= 64 else 32
    for name, lbo, sbo in (("LBO = k-block stride, SBO = mn-block stride", KS, MS), ("LBO = mn-block stride, SBO = k-block stride", MS, KS)):
        D = st.run(b, ctx, lib, a_img, b_img, [lbo, sbo, lbo, sbo, 2 * KS, 2 * KS, idesc_f16(128, N, 1, 1), (K // 16) | F16], ncols=ncols)
        n = min(N, ncols)
        D[-1, -1] = ref[-1, n - 1] if n == ncols else D[-1, -1]
        print(f"MN-major A and B, N={N:3d}, {name}: max|D-ref| = {np.abs(D[:, :n] - ref[:, :n]).max()}")
    # mixed: A MN-major, B K-major (the existing K-major image helper)
