mkdir -p gpurun_out
timeout 600 python tools/probe_backward_error.py 2>&1 | tail -3
timeout 1200 python tools/probe_gram_parity.py 0:0:1 1:0:1 2>&1 | tail -4
