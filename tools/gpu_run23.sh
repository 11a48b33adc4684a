mkdir -p gpurun_out
for sm in 32 64 128; do echo "== c5 per-patch, default workers, SVD_SMALL_MAX=$sm"; T4B_SVD_SMALL_MAX=$sm C5_NOPROF=1 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall | tail -2; done
