mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_c3_golden.py::test_c3_full_sweep_matches_oracle_golden tests/test_gpu_seam.py tests/test_gpu_svd.py::test_svd_reports_non_convergence -q -s 2>&1 | grep -v "^$" | cut -c1-400 > gpurun_out/pytest_r02c.log
tail -5 gpurun_out/pytest_r02c.log
