#!/usr/bin/env python
"""How the pivot sets of rrlu_fixtures.json were cross-checked: an exact-rational (fractions.Fraction) full-pivot
LU with the reference's selection rule - first maximum of |a_ij| in column-major scan order of the trailing
block (crates/tensor4all-core/src/matrixlu.rs:480-520 submatrix_argmax_col_major), stop on rank cap /
abs_tol / rel_tol exactly as rrlu_mut (:735-819).  Floating point never enters, so the expected row / column
index lists are independent of both the oracle and the CUDA kernel.

    python tests/golden/check_rrlu_fixtures.py        # exits non-zero on any mismatch
"""
import json
import os
import sys
from fractions import Fraction


def exact_rrlu(rows, max_rank=None, rel_tol=1e-14, abs_tol=0.0):
    a = [[Fraction(x) for x in r] for r in rows]
    m, n = len(a), len(a[0])
    rp, cp = list(range(m)), list(range(n))
    cap = min(m, n) if max_rank is None else min(m, n, max_rank)
    npiv, max_err = 0, Fraction(0)
    while npiv < cap:
        k = npiv
        best, bi, bj = Fraction(-1), k, k
        for j in range(k, n):               # column-major scan, strict '>' keeps the first maximum
            for i in range(k, m):
                v = abs(a[i][j])
                if v > best:
                    best, bi, bj = v, i, j
        # at least one pivot, then stop below the tolerances (:760-763); tiny-pivot guard (:768-779)
        if npiv > 0 and (best < Fraction(rel_tol) * max_err or best < Fraction(abs_tol)):
            break
        min_pivot = Fraction(0) if (rel_tol == 0.0 and abs_tol == 0.0) else Fraction(2.220446049250313e-16)
        if best <= min_pivot:
            break
        max_err = max(max_err, best)
        a[k], a[bi] = a[bi], a[k]
        rp[k], rp[bi] = rp[bi], rp[k]
        for r in a:
            r[k], r[bj] = r[bj], r[k]
        cp[k], cp[bj] = cp[bj], cp[k]
        piv = a[k][k]
        for i in range(k + 1, m):
            a[i][k] /= piv
            f = a[i][k]
            if f:
                for j in range(k + 1, n):
                    a[i][j] -= f * a[k][j]
        npiv += 1
    return npiv, rp[:npiv], cp[:npiv]


def main():
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rrlu_fixtures.json")))
    bad = 0
    for case in gold["cases"]:
        r, rows, cols = exact_rrlu(case["rows"], case.get("max_bond_dim"), case.get("rel_tol", 1e-14),
                                   case.get("abs_tol", 0.0))
        ok = (r == case["rank"] and rows == case["row_indices"] and cols == case["col_indices"])
        print(("ok   " if ok else "FAIL ") + case["name"], r, rows, cols)
        bad += 0 if ok else 1
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
