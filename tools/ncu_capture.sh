#!/bin/bash
# ncu evidence for profiles/: launch list of a window of one bench sweep + full captures of the hot kernels.
# Run under gpurun (one GPU).  Numbers printed by runs under ncu are never bench values.
#   bash tools/ncu_capture.sh <tag> [launch-skip] [launch-count]
set -x
mkdir -p gpurun_out
R=${1:-r01}
SKIP=${2:-0}
COUNT=${3:-37000}
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${SKIP} -c ${COUNT} --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_persistent -c 1 \
    -f -o gpurun_out/jacobi_2048_${R} python tools/probe_jac.py 2048x4096 > gpurun_out/ncu_jac_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_persistent -c 1 \
    -f -o gpurun_out/jacobi_512_${R} python tools/probe_jac.py 2048x512 > gpurun_out/ncu_jac512_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ws -c 6 \
    -f -o gpurun_out/gemm_ws_${R} python tools/probe_gemm.py > gpurun_out/ncu_gemm_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tsqr -s 20 -c 4 \
    -f -o gpurun_out/tsqr_${R} python tools/probe_prof.py qrr:4096x2048 > gpurun_out/ncu_tsqr_${R}.log 2>&1
ls -la gpurun_out
