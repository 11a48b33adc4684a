mkdir -p gpurun_out
for tx in 1 16 64 256; do echo "tolx=$tx"; T4B_JAC_TOLX=$tx timeout 300 python tools/probe_backward_error.py 2>&1 | tail -2 | cut -c1-700; done
timeout 1200 python tools/probe_gram_parity.py 0:0:1:1 0:0:1:16 0:0:1:64 0:0:1:256 2>&1 | tail -4
