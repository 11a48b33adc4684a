#!/bin/bash
# round-2 final build: launch list of one whole sweep + full captures of the Jacobi kernel and the blocked Cholesky
# diagonal kernel (the persistent Jacobi kernel is launched non-cooperatively under ncu: T4B_JAC_COOP=0)
set -x
mkdir -p gpurun_out
R=${1:-r02c}
T4B_JAC_COOP=0 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5 \
    > gpurun_out/bench_under_ncu_${R}.log 2>&1
T4B_JAC_COOP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_persistent -c 1 \
    -f -o gpurun_out/jacobi_2048_${R} python tools/probe_jac.py 2048x4096 > gpurun_out/ncu_jac_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_inv -s 2 -c 2 \
    -f -o gpurun_out/potrf_${R} python tools/probe_one_svd.py 2048 4096 > gpurun_out/ncu_potrf_${R}.log 2>&1
ls -la gpurun_out/*_${R}.ncu-rep
