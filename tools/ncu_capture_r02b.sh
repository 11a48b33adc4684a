#!/bin/bash
# second pass of the round-2 captures: the persistent Jacobi kernel must be launched non-cooperatively under ncu
# (T4B_JAC_COOP=0), and the batched one-CTA SVD is captured from a dedicated probe
set -x
mkdir -p gpurun_out
R=${1:-r02b}
T4B_JAC_COOP=0 timeout 1800 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5 \
    > gpurun_out/bench_under_ncu_${R}.log 2>&1
T4B_JAC_COOP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_persistent -c 1 \
    -f -o gpurun_out/jacobi_2048_${R} python tools/probe_jac.py 2048x4096 > gpurun_out/ncu_jac_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_inv -s 2 -c 2 \
    -f -o gpurun_out/potrf_${R} python tools/probe_one_svd.py 2048 4096 > gpurun_out/ncu_potrf_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svd_small_kernel -s 12 -c 1 \
    -f -o gpurun_out/svd_small_${R} python tools/probe_c1_batched.py > gpurun_out/ncu_svd_small_${R}.log 2>&1
ls -la gpurun_out/*_${R}.ncu-rep
