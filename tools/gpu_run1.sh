mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_r02.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_r02a.log
(timeout 300 python tools/probe_jac.py 2048x4096 2048x512 2>&1; T4B_JAC_OCC2=1 timeout 300 python tools/probe_jac.py 2048x4096 2048x512 2>&1; T4B_JAC_COOP=0 timeout 300 python tools/probe_jac.py 2048x4096 2>&1; T4B_JAC_CS=8 timeout 300 python tools/probe_jac.py 2048x512 2>&1) > gpurun_out/probe_jac_r02a.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
T4B_JAC_OCC2=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02a_occ2.json 2> gpurun_out/bench_r02a_occ2.err
tail -5 gpurun_out/pytest_r02a.log; cat gpurun_out/probe_jac_r02a.log | tail -40
