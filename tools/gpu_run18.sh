mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^\[t4b\]" | tail -60 > gpurun_out/pytest_r02r.log
grep -E "C3 saturated|C3 full sweep|C2 full|passed|failed|FAILED|Error" gpurun_out/pytest_r02r.log | tail -15
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02r.json 2> gpurun_out/bench_r02r.err
tail -c 600 gpurun_out/bench_r02r.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02r.json').read().strip().splitlines()[-1])
    print('c3', round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'])
    print('c5', json.dumps(d.get('c5')))
except Exception as e: print('ERR', e)
PY
