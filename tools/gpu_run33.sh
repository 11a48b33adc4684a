mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_svd.py tests/test_gpu_qr.py tests/test_gpu_seam.py tests/test_gpu_c3_golden.py tests/test_gpu_c2_c5_golden.py -m gpu -q -x 2>&1 | grep -v "^\[t4b\]" | tail -8
for v in new old; do if [ $v = old ]; then export T4B_CHOL_OLD=1; fi; timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02x_$v.json 2> gpurun_out/bench_r02x_$v.err; done
python - <<'PY'
import json
for v in ('new','old'):
    try:
        d=json.loads(open('gpurun_out/bench_r02x_%s.json'%v).read().strip().splitlines()[-1]); print('chol', v, round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], d['roofline']['frac'], {k:round(x,1) for k,x in d['kernel_profile_ms'].items() if x>15})
    except Exception as e: print('ERR', v, e)
PY
