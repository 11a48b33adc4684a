mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tci2.py tests/test_gpu_svd.py tests/test_gpu_simplett.py tests/test_gpu_seam.py tests/test_gpu_patches.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_r02h_new.log
tail -4 gpurun_out/pytest_r02h_new.log
timeout 300 python bench.py --workload c1 --steps 3 --warmup 2 > gpurun_out/bench_r02h_c1.json 2> gpurun_out/bench_r02h_c1.err
for l in 4 8 32; do T4B_SVD_LPP=$l timeout 300 python bench.py --workload c1 --steps 2 --warmup 1 > gpurun_out/bench_r02h_c1_lpp$l.json 2> gpurun_out/bench_r02h_c1_lpp$l.err; done
timeout 300 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_r02h_c5.json 2> gpurun_out/bench_r02h_c5.err
timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 > gpurun_out/bench_r02h_c2.json 2> gpurun_out/bench_r02h_c2.err
python - <<'PY'
import json
for f in ['c1','c1_lpp4','c1_lpp8','c1_lpp32','c5','c2']:
    try:
        d=json.loads(open('gpurun_out/bench_r02h_%s.json'%f).read().strip().splitlines()[-1])['record']
        if f.startswith('c1'): print(f, d['batch1']['ms_per_compress'], d['batched']['ms_per_batch'], d['batched']['kernel_profile_ms'])
        elif f=='c5': print(f, d['value'], d['phase_ms'])
        else: print(f, d)
    except Exception as e: print(f,'ERR',e)
PY
