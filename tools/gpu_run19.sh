mkdir -p gpurun_out
T4B_VERBOSE=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | tail -8
C5_N=32 timeout 300 python tools/probe_c5_batched.py 2>&1 | tail -5
T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | tail -5
