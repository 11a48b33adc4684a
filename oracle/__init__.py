"""ORACLE — test infrastructure only.

A CPU restatement (numpy / SciPy-LAPACK + a small C file for the bit-exact rrLU) of the reference
algorithms on the dense tensor-train hot path of tensor4all-rs.  Nothing in the product
(tensor4all-rs_b200/) imports, links or executes anything in this package: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, as the checker.

The genuine reference cannot be built in this environment (no Rust toolchain; its dense
arithmetic lives in the un-vendored tenferro-rs @ a21a4c6 crates), so the SVD / QR / GEMM calls
that the reference forwards to tenferro (faer) are served here by LAPACK gesdd / geqrf and BLAS
through NumPy/SciPy, and compared only through gauge-invariant quantities.  What the reference
implements in-tree (rrLU, all truncation-rank rules, every sweep schedule) is restated
line-by-line with file:line citations and pinned against the reference's own known answers
(tests/golden/, tests/test_oracle_golden.py): Hilbert rrLU ranks / last pivot errors, the
dense-kernel pivot fixtures, compute_retained_rank tables, zip-up == naive on the reference's
LCG MPOs, the two-scale compression fixture.
"""
