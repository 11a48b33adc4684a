mkdir -p gpurun_out
timeout 120 tensor4all-rs_b200/lib/probe_eig > gpurun_out/probe_eig.jsonl 2>&1; cat gpurun_out/probe_eig.jsonl
timeout 600 python tools/probe_backward_error.py 2>&1 | tail -4
