mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_svd.py tests/test_gpu_qr.py tests/test_gpu_seam.py tests/test_gpu_c3_golden.py tests/test_gpu_c2_c5_golden.py tests/test_gpu_treetn.py tests/test_gpu_simplett.py tests/test_gpu_fourier.py -m gpu -q -x -s 2>&1 | grep -v "^\[t4b\]" | tail -12
for v in 1 0; do echo "eig_v2=$v"; T4B_JAC_EIG_V2=$v T4B_VERBOSE=1 timeout 100 python tools/probe_one_svd.py 2048 4096 2>&1 | grep "jacobi nx" | tail -1; T4B_JAC_EIG_V2=$v T4B_VERBOSE=1 timeout 100 python tools/probe_one_svd.py 512 2048 2>&1 | grep "jacobi nx" | tail -1; done
