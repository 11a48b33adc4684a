mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensordot.py tests/test_gpu_qr.py tests/test_gpu_svd.py -q -x 2>&1 | tail -30 > gpurun_out/pytest_r02e_kernels.log
tail -3 gpurun_out/pytest_r02e_kernels.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_r02e.log
tail -8 gpurun_out/pytest_r02e.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err
T4B_GEMM_NOPERSIST=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02e_nopersist.json 2> gpurun_out/bench_r02e_nopersist.err
timeout 300 python tools/probe_gemm.py > gpurun_out/probe_gemm_r02e.json 2> gpurun_out/probe_gemm_r02e.err
T4B_GEMM_NOPERSIST=1 timeout 300 python tools/probe_gemm.py > gpurun_out/probe_gemm_r02e_nopersist.json 2> gpurun_out/probe_gemm_r02e_nopersist.err
python - <<'PY'
import json
for f in ['bench_r02e','bench_r02e_nopersist']:
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],1), d['roofline_contraction']['frac'], {k:v for k,v in list(d['kernel_profile_ms'].items())[:5]})
    except Exception as e: print(f,'ERR',e)
PY
tail -5 gpurun_out/probe_gemm_r02e.json gpurun_out/probe_gemm_r02e_nopersist.json
