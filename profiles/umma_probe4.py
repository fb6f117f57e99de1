"""Timing probe: cycles for R x 8 back-to-back tcgen05.mma kind::tf32 (M=128,N=64,K=8) — SS form vs TS form."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
b = g.load_package(); ctx = b.Context(0)
def image(mat, G_S, G_F, total_bytes):
    img = np.zeros(total_bytes // 4, np.float32)
    S, F = mat.shape
    s, f = np.meshgrid(np.arange(S), np.arange(F), indexing="ij")
    img[((s // 8) * G_S + (f // 4) * G_F + (s % 8) * 16 + (f % 4) * 4) // 4] = mat
    return img
def idesc(M, N): return (1 << 4) | (2 << 7) | (2 << 10) | ((N >> 3) << 17) | ((M >> 4) << 24)
def run(a_img, b_img, desc8, ncols=64):
    out = np.zeros((128, ncols), np.float32); d8 = np.array(desc8, np.uint32)
    b._lib.check(__import__("_selftest").load().b200rl_selftest_umma(ctx.h, a_img.ctypes.data_as(C.c_void_p), a_img.nbytes, b_img.ctypes.data_as(C.c_void_p), b_img.nbytes,
                                              d8.ctypes.data_as(C.c_void_p), ncols, out.ctypes.data_as(C.c_void_p)))
    return out
G_S, G_F = 2064, 128
A = np.ones((128, 64), np.float32); Bm = np.ones((64, 64), np.float32)
a_img = image(A, G_S, G_F, 16 * G_S + 64); b_img = image(Bm, G_S, G_F, 8 * G_S + 64)
for form, name in ((0, "SS"), (1, "TS")):
    for rep in (1, 3, 12, 48):
        kadv = 8 if form else 2 * G_F
        D = run(a_img, b_img, [G_F, G_S, G_F, G_S, kadv, 2 * G_F, idesc(128, 64), 8 | (form << 16) | (rep << 20)])
        cyc = D[127, 63]
        print(name, "MMAs", 8 * rep, "cycles", cyc, "per MMA", cyc / (8 * rep), "check", D[0, 0], "(expect", 64 * rep, ")")
for N in (64, 128, 256):
    pass
print("--- N scaling (SS form, 48 MMAs each) ---")
for N in (64, 128, 192, 256):
    Bn = np.ones((N, 64), np.float32)
    bi = image(Bn, G_S, G_F, (N // 8) * G_S + 64)
    D = run(a_img, bi, [G_F, G_S, G_F, G_S, 2 * G_F, 2 * G_F, idesc(128, N), 8 | (6 << 20)], ncols=64)
    print("N", N, "cycles/MMA", D[127, 63] / 48, "check", D[0, 0])
