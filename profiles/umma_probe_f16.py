"""Probe of tcgen05.mma kind::f16 (fp16 operands, K = 16 per instruction) for the 3-term fp16-split K7:
(1) SS form, K-major SWIZZLE_NONE images with 8-element (16-byte) core-matrix rows; (2) TS form, A packed two fp16 per TMEM
column; (3) pacing: cycles per instruction for the shapes K7 would issue, next to the kind::tf32 ones."""
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


def image16(mat, sbo, lbo, total_bytes):
    """K-major fp16 image: byte(r, k) = (r/8)*sbo + (k/8)*lbo + (r%8)*16 + (k%8)*2"""
    img = np.zeros(total_bytes // 2, np.float16)
    R, K = mat.shape
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    img[((r // 8) * sbo + (k // 8) * lbo + (r % 8) * 16 + (k % 8) * 2) // 2] = mat
    return img.view(np.float32)


def image32(mat, sbo, lbo, total_bytes):
    img = np.zeros(total_bytes // 4, np.float32)
    R, K = mat.shape
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    img[((r // 8) * sbo + (k // 4) * lbo + (r % 8) * 16 + (k % 4) * 4) // 4] = mat
    return img


def idesc_f16(M, N):
    return (1 << 4) | ((N >> 3) << 17) | ((M >> 4) << 24)


def idesc_tf32(M, N):
    return (1 << 4) | (2 << 7) | (2 << 10) | ((N >> 3) << 17) | ((M >> 4) << 24)


A = rng.integers(-4, 5, (128, 64)).astype(np.float32)
Bm = rng.integers(-4, 5, (128, 64)).astype(np.float32)      # N = 128 rows
LBO, SBO = 128, 1024                                        # 8 k-chunks of 8 elements per row group: 8 * 128 B
a_img = image16(A, SBO, LBO, 16 * SBO)
b_img = image16(Bm, SBO, LBO, 16 * SBO)
F16, TS, RAW = 1 << 17, 1 << 16, 1 << 18
for N in (64, 128):
    D = st.run(b, ctx, lib, a_img, b_img, [LBO, SBO, LBO, SBO, 2 * LBO, 2 * LBO, idesc_f16(128, N), 4 | F16], ncols=N)
    D[-1, -1] = (A @ Bm[:N].T)[-1, -1]
    print(f"f16 SS form N={N}: max|D-ref| =", np.abs(D - A @ Bm[:N].T).max())
# TS form: TMEM word [m][c] = fp16(A[m][2c]) | fp16(A[m][2c+1]) << 16  (hypothesis: low half = lower k)
h = A.astype(np.float16).view(np.uint16).astype(np.uint32)
raw = np.zeros((128, 64), np.uint32)
raw[:, :32] = h[:, 0::2] | (h[:, 1::2] << 16)
D = st.run(b, ctx, lib, raw.view(np.float32), b_img, [0, 0, LBO, SBO, 8, 2 * LBO, idesc_f16(128, 128), 4 | F16 | TS | RAW], ncols=128)
D[-1, -1] = (A @ Bm.T)[-1, -1]
print("f16 TS form (two fp16 per column, low half = lower k, 8 columns per k-step): max|D-ref| =", np.abs(D - A @ Bm.T).max())
raw2 = np.zeros((128, 64), np.uint32)
raw2[:, :32] = h[:, 1::2] | (h[:, 0::2] << 16)
D = st.run(b, ctx, lib, raw2.view(np.float32), b_img, [0, 0, LBO, SBO, 8, 2 * LBO, idesc_f16(128, 128), 4 | F16 | TS | RAW], ncols=128)
D[-1, -1] = (A @ Bm.T)[-1, -1]
print("f16 TS form (swapped halves): max|D-ref| =", np.abs(D - A @ Bm.T).max())
# feature-major GEMM3-like operands with the padded strides K7 uses (K = 128 samples): row stride between 8-sample chunks
P = rng.integers(-3, 4, (128, 128)).astype(np.float32)   # A rows = 128 (dP2^T full|lo stacked), K = 128 samples
Hh = rng.integers(-3, 4, (128, 128)).astype(np.float32)  # B rows = 128 (H1^T full|lo stacked), K = 128 samples
for lbo in (128, 144, 136):
    sbo = 16 * lbo + 16
    pa = image16(P, sbo, lbo, 16 * sbo + 64); hb = image16(Hh, sbo, lbo, 16 * sbo + 64)
    D = st.run(b, ctx, lib, pa, hb, [lbo, sbo, lbo, sbo, 2 * lbo, 2 * lbo, idesc_f16(128, 128), 8 | F16], ncols=128)
    D[-1, -1] = (P @ Hh.T)[-1, -1]
    print(f"f16 SS GEMM3-like (M=128, N=128, K=128, LBO={lbo}, SBO={sbo}): max|D-ref| =", np.abs(D - P @ Hh.T).max())

# ---- pacing -------------------------------------------------------------------------------------------------------------
def cycles(desc8, a, bb, ncols, reps=64):
    d = list(desc8)
    base = st.run(b, ctx, lib, a, bb, d[:7] + [d[7] | (1 << 20)], ncols=ncols)[-1, -1]
    many = st.run(b, ctx, lib, a, bb, d[:7] + [d[7] | (reps << 20)], ncols=ncols)[-1, -1]
    n = d[7] & 0xFFFF
    return (many - base) / ((reps - 1) * n)

a32 = image32(A, 2064, 128, 16 * 2064 + 64); b32 = image32(Bm, 2064, 128, 16 * 2064 + 64)
print("cycles per tcgen05.mma (steady state, one issuing thread, dependent accumulates):")
for N in (64, 128):
    print(f"  tf32 TS M=128 N={N:3d} K=8 :", round(cycles([128, 2064, 128, 2064, 8, 256, idesc_tf32(128, N), 8 | TS], a32, b32, N), 1))
    print(f"  tf32 SS M=128 N={N:3d} K=8 :", round(cycles([128, 2064, 128, 2064, 256, 256, idesc_tf32(128, N), 8], a32, b32, N), 1))
    print(f"  f16  TS M=128 N={N:3d} K=16:", round(cycles([0, 0, LBO, SBO, 8, 2 * LBO, idesc_f16(128, N), 4 | F16 | TS | RAW], raw.view(np.float32), b_img, N), 1))
    print(f"  f16  SS M=128 N={N:3d} K=16:", round(cycles([LBO, SBO, LBO, SBO, 2 * LBO, 2 * LBO, idesc_f16(128, N), 4 | F16], a_img, b_img, N), 1))
