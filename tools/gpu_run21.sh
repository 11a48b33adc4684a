mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_qr.py tests/test_gpu_seam.py tests/test_gpu_linalg_extra.py -m gpu -q -x 2>&1 | tail -4
echo "== c5 per-patch, 4 workers"; C5_NOPROF=1 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
echo "== c5 per-patch, 8 workers"; C5_NOPROF=1 T4B_PATCH_WORKERS=8 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
echo "== c5 per-patch, 1 worker"; C5_NOPROF=1 C5_REPS=3 T4B_PATCH_WORKERS=1 T4B_PATCH_BATCHED=0 timeout 300 python tools/probe_c5_batched.py 2>&1 | grep -v "^\[t4b\]" | grep wall
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02s.json 2> gpurun_out/bench_r02s.err
timeout 300 python bench.py --workload c1 --steps 3 --warmup 2 > gpurun_out/bench_r02s_c1.json 2> gpurun_out/bench_r02s_c1.err
python - <<'PY'
import json
def rec(f): return json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
try:
    d=rec('bench_r02s'); print('c3', round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], {k:round(v,1) for k,v in d['kernel_profile_ms'].items() if v>20})
except Exception as e: print('c3 ERR', e)
try:
    d=rec('bench_r02s_c1'); print('c1', json.dumps(d['record'])[:700])
except Exception as e: print('c1 ERR', e)
PY
