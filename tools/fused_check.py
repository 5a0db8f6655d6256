"""Diagnostic run of the single-launch tier (lu_fused.cu) against the oracle: per shape, pivots / info / bit equality,
and where the first difference sits. Then timings of C3 (n = 128, 50k) with the tier on and off.
   python tools/fused_check.py [quick]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from magma_b200 import batched as mb

torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)

def run(A0, m, ldda=None, offset=0):
    batch, n, _ = A0.shape
    ldda = m if ldda is None else ldda
    db = mb.DeviceBatch(batch, m, n, ldda=ldda, queue=q)
    Ain = np.full((batch, n, ldda), 7.25)
    Ain[:, :, :m] = A0[:, :, :m]
    db.upload(Ain)
    rc = db.getrf()
    LU, ipiv, info = db.download()
    pad_ok = bool(np.all(LU[:, :, m:] == 7.25))
    return rc, LU[:, :, :m].copy(), ipiv, info, pad_ok

def check(m, n, batch, ldda=None, A0=None, tag=""):
    if A0 is None:
        A0, _ = oracle.random_batch(batch, m, n)
    t0 = time.time()
    rc, LU, ipiv, info, pad_ok = run(A0, m, ldda)
    ref = np.ascontiguousarray(A0[:, :, :m]).copy()
    ipr, infr = oracle.getrf_batched(ref, m)
    okp = np.array_equal(ipiv, ipr); oki = np.array_equal(info, infr); okl = np.array_equal(LU, ref)
    msg = f"m={m} n={n} ld={ldda} b={A0.shape[0]} {tag}: rc={rc} piv={'ok' if okp else 'BAD'} info={'ok' if oki else 'BAD'} LU={'bit' if okl else 'DIFF'} pad={'ok' if pad_ok else 'BAD'}"
    if not okp:
        bb, kk = np.argwhere(ipiv != ipr)[0]
        msg += f" | first piv diff: matrix {bb} step {kk}: got {ipiv[bb, kk]} want {ipr[bb, kk]} (bad matrices {np.any(ipiv != ipr, axis=1).sum()})"
    if not okl:
        d = np.argwhere(LU != ref)
        # LU[b, col, row]
        cols = np.unique(d[:, 1]); rows = np.unique(d[:, 2])
        bb, cc, rr = d[0]
        msg += f" | LU diffs {len(d)}: first matrix {bb} col {cc} row {rr} got {LU[bb, cc, rr]:.6g} want {ref[bb, cc, rr]:.6g}; cols {cols.min()}..{cols.max()} rows {rows.min()}..{rows.max()} maxabs {np.nanmax(np.abs(LU - ref)):.3g}"
        # per 8x8 block map of the first bad matrix
        bad = (LU[bb] != ref[bb])  # [col, row]
        nb_c, nb_r = (n + 7) // 8, (m + 7) // 8
        lines = []
        for rb in range(nb_r):
            lines.append("".join("X" if bad[8 * cb:8 * cb + 8, 8 * rb:8 * rb + 8].any() else "." for cb in range(nb_c)))
        msg += "\n   block map (rows down, cols across):\n   " + "\n   ".join(lines)
    print(msg, flush=True)
    return okp and oki and okl and pad_ok and rc == 0

quick = len(sys.argv) > 1
mb.set_fused_max(128)
print("rcp selftest mismatches:", mb.rcp_selftest(200_000_000, q), flush=True)
shapes = [(128, 128, 6), (64, 64, 6), (40, 40, 8), (128, 64, 4), (64, 128, 4), (96, 96, 4), (100, 100, 4), (128, 72, 3), (72, 128, 3),
          (33, 33, 5), (65, 65, 5), (127, 127, 3), (128, 127, 3), (127, 128, 3), (97, 113, 3), (113, 97, 3), (48, 120, 3), (120, 48, 3),
          (128, 8, 3), (8, 128, 3), (128, 1, 2), (1, 128, 2), (70, 66, 3), (66, 70, 3), (128, 96, 3), (96, 128, 3), (128, 104, 2)]
allok = True
for (m, n, b) in shapes:
    allok &= check(m, n, b)
# padded / odd leading dimensions (generic load path when ld is odd)
for (m, n, ld) in [(128, 128, 130), (128, 128, 129), (100, 100, 101), (64, 64, 72), (127, 127, 128), (90, 90, 90)]:
    allok &= check(m, n, 3, ldda=ld)
# structured: zero matrix, ones, identity, flipped identity, integer ties, zero column
n = 128
rng = np.random.default_rng(0)
mats = [np.zeros((n, n)), np.ones((n, n)), np.eye(n), np.fliplr(np.eye(n)), rng.integers(-3, 4, size=(n, n)).astype(float)]
Z = rng.random((n, n)); Z[:, 70] = 0.0; mats.append(Z)
Z2 = rng.random((n, n)); Z2[5, :] = 0.0; mats.append(Z2)
allok &= check(n, n, len(mats), A0=np.stack(mats), tag="structured")
print("ALL OK" if allok else "FAILURES", flush=True)

def flops(n):
    return 0.5 * n * (n * (n - n / 3.0 - 1.0) + n) + 2.0 * n / 3.0 + 0.5 * n * (n * (n - n / 3.0) - n) + n / 6.0

def timeit(n, batch, reps=4):
    db = mb.DeviceBatch(batch, n, n, queue=q)
    seed = np.array([0, 0, 0, 1], dtype=np.int32)
    mb.dlarnv_uniform(seed, batch * n * n, db.A, q); q.sync()
    A0 = db.A.clone(); ts = []
    for _ in range(reps):
        db.A.copy_(A0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); db.getrf(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    del db, A0; torch.cuda.empty_cache()
    return min(ts)

if not quick:
    for (n, batch) in [(128, 50000), (96, 50000), (64, 100000), (48, 100000), (40, 200000)]:
        mb.set_fused_max(128); t1 = timeit(n, batch)
        mb.set_fused_max(0); t0 = timeit(n, batch)
        mb.set_fused_max(128)
        print(f"n={n} batch={batch}: fused {t1:.3f} ms ({flops(n) * batch / t1 / 1e6:.0f} GF/s)   previous path {t0:.3f} ms", flush=True)
