mkdir -p gpurun_out
T4B_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_qr.py -q -x 2>&1 | grep -v "^\[t4b\] gemm\|jacobi nx" | tail -30 > gpurun_out/pytest_r02j_svd.log
tail -14 gpurun_out/pytest_r02j_svd.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_r02j.log
tail -8 gpurun_out/pytest_r02j.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err
T4B_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5 2>&1 | grep "Cholesky" | awk '{print $2,$3,$4,$5,$6,$7,$8}' | sort | uniq -c | sort -rn | head -12 > gpurun_out/gram_decisions_r02j.txt
cat gpurun_out/gram_decisions_r02j.txt
python - <<'PY'
import json
for f in ['bench_r02j']:
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], d['kernel_profile_ms'], d['result'])
    except Exception as e: print(f,'ERR',e)
PY
