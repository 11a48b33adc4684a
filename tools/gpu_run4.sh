mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensordot.py tests/test_gpu_qr.py tests/test_gpu_svd.py -q -x 2>&1 | tail -30 > gpurun_out/pytest_r02d_kernels.log
tail -3 gpurun_out/pytest_r02d_kernels.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_r02d.log
tail -8 gpurun_out/pytest_r02d.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err
T4B_QR_LEAF_OLD=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02d_oldleaf.json 2> gpurun_out/bench_r02d_oldleaf.err
T4B_GEMM_NOPERSIST=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02d_nopersist.json 2> gpurun_out/bench_r02d_nopersist.err
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 > gpurun_out/bench_r02d_c4.json 2> gpurun_out/bench_r02d_c4.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:rrlu_kernel -c 1 -o gpurun_out/rrlu_r02 -f python bench.py --workload c4 --steps 1 --warmup 0 > gpurun_out/ncu_rrlu_r02.log 2>&1
python - <<'PY'
import json
for f in ['bench_r02d','bench_r02d_oldleaf','bench_r02d_nopersist']:
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],1), d['roofline_contraction']['frac'], {k:v for k,v in list(d['kernel_profile_ms'].items())[:5]})
    except Exception as e: print(f,'ERR',e)
PY
