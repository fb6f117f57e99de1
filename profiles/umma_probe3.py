"""Probe: (1) A operand sourced from TMEM (.ts form), (2) TMEM lane mapping of an M = 64 accumulator,
(3) M = 128 MMA whose A rows 64..127 are junk (only lanes 0..63 meaningful)."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)
rng = np.random.default_rng(0)
def image(mat, G_S, G_F, total_bytes):
    img = np.zeros(total_bytes // 4, np.float32)
    S, F = mat.shape
    s, f = np.meshgrid(np.arange(S), np.arange(F), indexing="ij")
    img[((s // 8) * G_S + (f // 4) * G_F + (s % 8) * 16 + (f % 4) * 4) // 4] = mat
    return img
def idesc(M, N, a_mn=0, b_mn=0):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)
def run(a_img, b_img, desc8, ncols=64):
    out = np.zeros((128, ncols), np.float32)
    d8 = np.array(desc8, np.uint32)
    b._lib.check(__import__("_selftest").load().b200rl_selftest_umma(ctx.h, a_img.ctypes.data_as(C.c_void_p), a_img.nbytes, b_img.ctypes.data_as(C.c_void_p), b_img.nbytes,
                                              d8.ctypes.data_as(C.c_void_p), ncols, out.ctypes.data_as(C.c_void_p)))
    return out
G_S, G_F = 2064, 128
A = rng.integers(-4, 5, (128, 64)).astype(np.float32); Bm = rng.integers(-4, 5, (64, 64)).astype(np.float32)
a_img = image(A, G_S, G_F, 16 * G_S + 64); b_img = image(Bm, G_S, G_F, 8 * G_S + 64)
ref = A @ Bm.T
D = run(a_img, b_img, [G_F, G_S, G_F, G_S, 8, 2 * G_F, idesc(128, 64), 8 | (1 << 16)])
print("TS form (A from TMEM, 8 columns per k-step): max|D-ref| =", np.abs(D - ref).max())
# M = 64: where do rows land?
A64 = A[:64]
D = run(image(A64, G_S, G_F, 16 * G_S + 64), b_img, [G_F, G_S, G_F, G_S, 2 * G_F, 2 * G_F, idesc(64, 64), 8])
ref64 = A64 @ Bm.T
lanes = {r: [l for l in range(128) if np.array_equal(D[l], ref64[r])] for r in range(64)}
print("M=64 row->lane:", [(r, lanes[r]) for r in (0, 1, 15, 16, 17, 31, 32, 33, 47, 48, 63)])
# f-major operands for the weight-gradient GEMM: A3 = P^T (M = 64 features padded to 128 rows, K = 128 samples), B3 = H^T (N = 64, K = 128)
P = rng.integers(-3, 4, (128, 64)).astype(np.float32); Hh = rng.integers(-3, 4, (128, 64)).astype(np.float32)
GF2 = 128; GS2 = 32 * GF2 + 16      # rows = features (8 groups), K chunks = 32 (4 samples each)
pt_img = image(P.T.copy(), GS2, GF2, 16 * GS2 + 64)      # only 8 of 16 row groups are real
ht_img = image(Hh.T.copy(), GS2, GF2, 8 * GS2 + 64)
refw = P.T @ Hh
D = run(pt_img, ht_img, [GF2, GS2, GF2, GS2, 2 * GF2, 2 * GF2, idesc(128, 64), 16])
print("f-major GEMM3 (M=128 padded, K=128): max|D[:64]-ref| =", np.abs(D[:64] - refw).max())
D = run(pt_img, ht_img, [GF2, GS2, GF2, GS2, 2 * GF2, 2 * GF2, idesc(64, 64), 16])
print("f-major GEMM3 (M=64): rows found at lanes", [(r, [l for l in range(128) if np.array_equal(D[l], refw[r])]) for r in (0, 15, 16, 32, 63)])
