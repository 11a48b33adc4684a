mkdir -p gpurun_out
echo "== default (4 workers, coop)"; T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
echo "== coop=0, 4 workers"; T4B_JAC_COOP=0 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
echo "== coop=0, 12 workers"; T4B_JAC_COOP=0 T4B_PATCH_WORKERS=12 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
echo "== coop=1, 12 workers"; T4B_PATCH_WORKERS=12 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
echo "== nogram, 8 workers"; T4B_SVD_NOGRAM=1 T4B_PATCH_WORKERS=8 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
