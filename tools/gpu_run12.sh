mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_qr.py tests/test_gpu_treetn.py tests/test_gpu_tree.py tests/test_gpu_c3_golden.py tests/test_gpu_c2_c5_golden.py -q 2>&1 | tail -12 > gpurun_out/pytest_r02l.log
tail -8 gpurun_out/pytest_r02l.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02l.json 2> gpurun_out/bench_r02l.err
python - <<'PY'
import json
for f in ['bench_r02l']:
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], d['kernel_profile_ms'], d['result'])
    except Exception as e: print(f,'ERR',e)
PY
bash tools/ncu_capture_r02b.sh r02b > gpurun_out/ncu_capture_r02b.log 2>&1
tail -6 gpurun_out/ncu_capture_r02b.log
