mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_treetci.py tests/test_partitioned_contract.py -q -x 2>&1 | tail -30 > gpurun_out/pytest_r02f_new.log
tail -5 gpurun_out/pytest_r02f_new.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_r02f.log
tail -8 gpurun_out/pytest_r02f.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
timeout 300 python tools/probe_gemm.py > gpurun_out/probe_gemm_r02f.json 2> gpurun_out/probe_gemm_r02f.err
timeout 300 python bench.py --workload c1 --steps 3 --warmup 2 > gpurun_out/bench_r02f_c1.json 2> gpurun_out/bench_r02f_c1.err
timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 > gpurun_out/bench_r02f_c2.json 2> gpurun_out/bench_r02f_c2.err
timeout 300 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_r02f_c5.json 2> gpurun_out/bench_r02f_c5.err
python - <<'PY'
import json
for f in ['bench_r02f']:
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],1), d['roofline_contraction']['frac'], {k:v for k,v in list(d['kernel_profile_ms'].items())[:6]}, d.get('c5',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
for f in ['c1','c2','c5']:
    try: print(open('gpurun_out/bench_r02f_%s.json'%f).read()[:600])
    except Exception as e: print(f,'ERR',e)
PY
