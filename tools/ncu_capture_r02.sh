#!/bin/bash
# Round-2 ncu evidence: launch list of one whole bench sweep + full captures of the kernels that changed or were
# never profiled (Jacobi on the Cholesky factor, potrf + inverse, prrLU, single-CTA SVD, stream-K GEMM).
#   gpurun -- bash tools/ncu_capture_r02.sh r02
set -x
mkdir -p gpurun_out
R=${1:-r02}
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5 \
    > gpurun_out/bench_under_ncu_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_persistent -c 1 \
    -f -o gpurun_out/jacobi_2048_${R} python tools/probe_jac.py 2048x4096 > gpurun_out/ncu_jac_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_inv -s 2 -c 2 \
    -f -o gpurun_out/potrf_${R} python tools/probe_one_svd.py 2048 4096 > gpurun_out/ncu_potrf_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rrlu_kernel -c 2 \
    -f -o gpurun_out/rrlu_${R} python bench.py --workload c4 --steps 1 --warmup 0 > gpurun_out/ncu_rrlu_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svd_small_kernel -s 30 -c 1 \
    -f -o gpurun_out/svd_small_${R} python bench.py --workload c1 --steps 1 --warmup 0 > gpurun_out/ncu_svd_small_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ws -c 6 \
    -f -o gpurun_out/gemm_ws_${R} python tools/probe_gemm.py > gpurun_out/ncu_gemm_${R}.log 2>&1
ls -la gpurun_out | tail -20
