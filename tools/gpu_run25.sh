mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_seam.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02u.json 2> gpurun_out/bench_r02u.err
T4B_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5 2> gpurun_out/verbose_r02u.err > /dev/null
grep "nx=2048" gpurun_out/verbose_r02u.err | sed -n '3p;10p'
grep "nx=512 npad=512 .*sweeps=10" gpurun_out/verbose_r02u.err | sed -n '3p'
python - <<'PY'
import json
def rec(f): return json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
try:
    d=rec('bench_r02u'); print('c3', round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], d['roofline']['frac'], {k:round(v,1) for k,v in d['kernel_profile_ms'].items() if v>20})
except Exception as e: print('c3 ERR', e)
PY
