mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_qr.py tests/test_gpu_seam.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python tools/probe_backward_error.py 2>&1 | tail -3
timeout 900 python tools/probe_gram_parity.py 0 1 2>&1 | tail -3
