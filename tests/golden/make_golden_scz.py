"""Regenerate tests/golden/lu_golden_scz.npz: LAPACK {s,c,z}getrf / getrs outputs (scipy.linalg.lapack, OpenBLAS 0.3.31.dev
bundled with scipy -- third-party, not under /root/reference) for a few shapes in each of the other three precisions.
Run from the repo root:

    python tests/golden/make_golden_scz.py

Inputs: numpy default_rng(seed) uniform (0,1) (both parts for c / z) cast to the precision. Stored per case: input A,
right-hand sides B, LAPACK's pivots, factors and solutions for trans = N / T / C. tests/test_oracle.py checks the
per-precision oracle (oracle/lu_oracle_scz.c) against them: pivots identical, values to a few n * eps.
"""
import os

import numpy as np
from scipy.linalg import lapack

DT = {"s": np.float32, "c": np.complex64, "z": np.complex128}
SHAPES = [(4, 4), (16, 16), (32, 32), (64, 64), (32, 12), (12, 32), (100, 70)]  # (m, n)


def main():
    out = {}
    for p, dt in DT.items():
        getrf = getattr(lapack, p + "getrf")
        getrs = getattr(lapack, p + "getrs")
        for (m, n) in SHAPES:
            rng = np.random.default_rng(1000 * m + n)
            A = rng.random((m, n))
            if p != "s":
                A = A + 1j * rng.random((m, n))
            A = A.astype(dt)
            lu, piv, info = getrf(A)
            assert info == 0
            key = f"{p}_m{m}_n{n}"
            out[key + "_A"] = A
            out[key + "_LU"] = lu
            out[key + "_ipiv"] = (piv + 1).astype(np.int32)
            if m == n:
                B = rng.random((n, 2))
                if p != "s":
                    B = B + 1j * rng.random((n, 2))
                B = B.astype(dt)
                out[key + "_B"] = B
                for t, name in ((0, "N"), (1, "T"), (2, "C")):
                    x, info = getrs(lu, piv, B, trans=t)
                    assert info == 0
                    out[key + "_X" + name] = x
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lu_golden_scz.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
