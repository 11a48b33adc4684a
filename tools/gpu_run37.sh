mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^\[t4b\]" | tail -6 > gpurun_out/pytest_r02_final.log; tail -3 gpurun_out/pytest_r02_final.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -c 300 gpurun_out/bench_r02_final.err
timeout 120 tensor4all-rs_b200/lib/probe_eig > gpurun_out/probe_eig_r02.jsonl 2>&1
timeout 60 tensor4all-rs_b200/lib/probe_lat > gpurun_out/probe_lat_r02.txt 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_final.json').read().strip().splitlines()[-1]); print('c3', round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], d['roofline']['frac'], d['gpu_launches'], {k:round(x,1) for k,x in d['kernel_profile_ms'].items() if x>10}); c=d.get('c5') or {}; print('c5', c.get('value'), c.get('ms'), c.get('phase_ms')); print('cpu', json.dumps(d.get('cpu_baseline'))[:300]); print('clocks', d.get('clocks'))
PY
