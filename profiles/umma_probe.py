"""Probe the tcgen05 operand-layout / TMEM mapping conventions (run under gpurun)."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)
rng = np.random.default_rng(0)

def image(mat, G_S, G_F, total_bytes):
    """mat[s, f] -> bytes at (s//8)*G_S + (f//4)*G_F + (s%8)*16 + (f%4)*4"""
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

def report(name, D, ref_rows):
    """ref_rows: dict row_index -> expected row vector; find which TMEM lane holds each"""
    print("==", name)
    ok = 0
    lanes = {}
    for r, vec in ref_rows.items():
        hit = [l for l in range(128) if np.array_equal(D[l, :len(vec)], vec)]
        lanes[r] = hit
        ok += len(hit) > 0
    print("  rows found:", ok, "/", len(ref_rows))
    items = sorted(lanes.items())
    print("  row->lane (first 8):", items[:8])
    print("  row->lane (rows 16,17,32,48,63):", [(r, lanes.get(r)) for r in (16, 17, 32, 48, 63)])
    ident = all(lanes[r] == [r] for r in lanes)
    print("  identity mapping:", ident)
    return lanes

G_S, G_F = 2064, 128
# ---- S1: K-major A (128 x 64), K-major B (64 x 64) --------------------------------------------
A = rng.integers(-4, 5, (128, 64)).astype(np.float32)
Bm = rng.integers(-4, 5, (64, 64)).astype(np.float32)      # B[n, k]
a_img = image(A, G_S, G_F, 16 * G_S + 64)
b_img = image(Bm, G_S, G_F, 8 * G_S + 64)
D = run(a_img, b_img, [G_F, G_S, G_F, G_S, 2 * G_F, 2 * G_F, idesc(128, 64, 0, 0), 8])
ref = A @ Bm.T
report("S1 K-major A(128x64) x K-major B(64x64), M=128", D, {r: ref[r] for r in range(128)})
print("  max |D-ref|:", np.abs(D - ref).max())
# the same with the alternative reading (LBO/SBO swapped) to be sure which is which
D2 = run(a_img, b_img, [G_S, G_F, G_S, G_F, 2 * G_F, 2 * G_F, idesc(128, 64, 0, 0), 8])
print("  swapped LBO/SBO max |D-ref|:", np.abs(D2 - ref).max())

# ---- S2: MN-major A (M = 64 features, K = 128 samples), MN-major B (N = 64, K = 128) -----------
P = rng.integers(-3, 4, (128, 64)).astype(np.float32)      # dP2[s, j]
Hh = rng.integers(-3, 4, (128, 64)).astype(np.float32)     # H1[s, i]
p_img = image(P, G_S, G_F, 16 * G_S + 4096 + 64)
h_img = image(Hh, G_S, G_F, 16 * G_S + 4096 + 64)
ref = P.T @ Hh                                             # dW2[j, i]
D = run(p_img, h_img, [G_S, G_F, G_S, G_F, G_S, G_S, idesc(64, 64, 1, 1), 16])
report("S2 MN-major A(64 x K128) x MN-major B(64 x K128), M=64", D, {r: ref[r] for r in range(64)})
D = run(p_img, h_img, [G_F, G_S, G_F, G_S, G_S, G_S, idesc(64, 64, 1, 1), 16])
report("S2' same, LBO/SBO swapped", D, {r: ref[r] for r in range(64)})
# ---- S3: same with M = 128 (rows 64..127 read past the feature range) --------------------------
D = run(p_img, h_img, [G_S, G_F, G_S, G_F, G_S, G_S, idesc(128, 64, 1, 1), 16])
report("S3 MN-major, M=128 (only rows 0..63 meaningful)", D, {r: ref[r] for r in range(64)})
# ---- S4: 3xTF32 accuracy: random fp32 data, hi/lo split, three products ------------------------
A = rng.standard_normal((128, 64)).astype(np.float32); Bm = rng.standard_normal((64, 64)).astype(np.float32)
def split(x):
    hi = (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi, (x - hi).astype(np.float32)
Ah, Al = split(A); Bh, Bl = split(Bm)
ref = A.astype(np.float64) @ Bm.astype(np.float64).T
D1 = run(image(A, G_S, G_F, 16 * G_S + 64), image(Bm, G_S, G_F, 8 * G_S + 64), [G_F, G_S, G_F, G_S, 2 * G_F, 2 * G_F, idesc(128, 64, 0, 0), 8])
print("== S4 plain tf32 rel err:", np.linalg.norm(D1 - ref) / np.linalg.norm(ref))
tot = np.zeros_like(ref)
for x, y in ((Ah, Bh), (Ah, Bl), (Al, Bh)):
    tot += run(image(x, G_S, G_F, 16 * G_S + 64), image(y, G_S, G_F, 8 * G_S + 64), [G_F, G_S, G_F, G_S, 2 * G_F, 2 * G_F, idesc(128, 64, 0, 0), 8])
print("== S4 3xTF32 (summed on host) rel err:", np.linalg.norm(tot - ref) / np.linalg.norm(ref))
