mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_c3_golden.py -q -k "saturated" 2>&1 | tail -5
T4B_SVD_NOGRAM=1 timeout 600 python -m pytest tests/test_gpu_c3_golden.py -q -k "saturated" 2>&1 | tail -3
T4B_GEMM_NOPERSIST=1 timeout 600 python -m pytest tests/test_gpu_c3_golden.py -q -k "saturated" 2>&1 | tail -3
T4B_VERBOSE=1 timeout 600 python -m pytest tests/test_gpu_c3_golden.py -q -k "saturated" 2>&1 | grep "Cholesky" | sed 's/\[t4b\] //' | sort | uniq -c | sort -rn | head -40
