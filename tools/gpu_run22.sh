mkdir -p gpurun_out
nproc
timeout 900 python -m pytest tests/test_gpu_patches.py tests/test_partitioned_contract.py tests/test_gpu_treetn.py -m gpu -q -x 2>&1 | tail -4
echo "== c5 per-patch, 4 workers"; C5_NOPROF=1 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
echo "== c5 per-patch, 8 workers"; C5_NOPROF=1 T4B_PATCH_WORKERS=8 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
echo "== c5 per-patch, 16 workers"; C5_NOPROF=1 T4B_PATCH_WORKERS=16 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
