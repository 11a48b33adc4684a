mkdir -p gpurun_out
T4B_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_svd.py -q -x 2>&1 | grep -v "^\[t4b\] gemm" | tail -25 > gpurun_out/pytest_r02i_svd.log
tail -12 gpurun_out/pytest_r02i_svd.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_r02i.log
tail -8 gpurun_out/pytest_r02i.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err
T4B_SVD_NOGRAM=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02i_nogram.json 2> gpurun_out/bench_r02i_nogram.err
T4B_VERBOSE=1 timeout 300 python tools/probe_one_svd.py 4096 2048 2>&1 | grep -i "gram\|chol" | head -5
python - <<'PY'
import json
for f in ['bench_r02i','bench_r02i_nogram']:
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],1), d['roofline_contraction']['frac'], d['kernel_profile_ms'], d['result'])
    except Exception as e: print(f,'ERR',e)
PY
