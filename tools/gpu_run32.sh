mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^\[t4b\]" | tail -6 > gpurun_out/pytest_r02w.log; tail -4 gpurun_out/pytest_r02w.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r02w.json 2> gpurun_out/bench_r02w.err; tail -c 600 gpurun_out/bench_r02w.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02w.json').read().strip().splitlines()[-1]); print('c3', round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], d['roofline']['frac'], {k:round(x,1) for k,x in d['kernel_profile_ms'].items() if x>10}); print('c5', json.dumps(d.get('c5'))[:900]); print('cpu', json.dumps(d.get('cpu_baseline'))[:500])
PY
