"""Decode how tcgen05.mma addresses an MN-major SWIZZLE_NONE operand: fill the operand image with
word-index codes, multiply by a one-hot K-major operand, read the fetched word index out of D."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)

def image(mat, G_S, G_F, total_bytes):
    img = np.zeros(total_bytes // 4, np.float32)
    S, F = mat.shape
    s, f = np.meshgrid(np.arange(S), np.arange(F), indexing="ij")
    off = (s // 8) * G_S + (f // 4) * G_F + (s % 8) * 16 + (f % 4) * 4
    img[off // 4] = mat
    return img

def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)

def run(a_img, b_img, desc8, ncols=64):
    out = np.zeros((128, ncols), np.float32)
    d8 = np.array(desc8, np.uint32)
    b._lib.check(__import__("_selftest").load().b200rl_selftest_umma(ctx.h, a_img.ctypes.data_as(C.c_void_p), a_img.nbytes, b_img.ctypes.data_as(C.c_void_p), b_img.nbytes,
                                              d8.ctypes.data_as(C.c_void_p), ncols, out.ctypes.data_as(C.c_void_p)))
    return out

NW = 12288  # words in the coded image (48 KB)
codes_lo = (np.arange(NW) % 2048).astype(np.float32)
codes_hi = (np.arange(NW) // 2048).astype(np.float32)
onehot_B = np.zeros((64, 8), np.float32); onehot_B[np.arange(8), np.arange(8)] = 1.0       # B[n,k] = (n==k)
b_img = image(onehot_B, 2064, 128, 8 * 2064 + 64)

def decode_A(a_major, lbo, sbo, M=128, label=""):
    D_lo = run(codes_lo, b_img, [lbo, sbo, 128, 2064, 0, 0, idesc(M, 64, a_major, 0), 1])
    D_hi = run(codes_hi, b_img, [lbo, sbo, 128, 2064, 0, 0, idesc(M, 64, a_major, 0), 1])
    W = (D_hi[:, :8] * 2048 + D_lo[:, :8]).astype(np.int64)     # word index fetched for A(m, k)
    print("==", label, "A major", a_major, "LBO", lbo, "SBO", sbo, "M", M)
    for m in (0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 32, 63, 64, 127):
        print("   m=%3d  byte offsets for k=0..7:" % m, (W[m] * 4).tolist())
    return W

decode_A(0, 128, 2064, label="control K-major")
decode_A(1, 2064, 128, label="MN-major (LBO=k-group stride?, SBO=mn-group stride?)")
decode_A(1, 128, 2064, label="MN-major swapped")
decode_A(1, 4096, 256, label="MN-major distinct strides")
decode_A(1, 2064, 128, M=64, label="MN-major M=64")
