mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_patches.py tests/test_partitioned_contract.py tests/test_gpu_svd.py tests/test_gpu_c2_c5_golden.py -m gpu -q -x 2>&1 | grep -v "^\[t4b\]" | tail -4
timeout 300 python bench.py --workload c5 --steps 2 --warmup 1 > gpurun_out/bench_r02z_c5.json 2> gpurun_out/bench_r02z_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02z_c5.json').read().strip().splitlines()[-1]); r=d.get('record',d)
print('c5', r.get('value'), r.get('ms'), r.get('phase_ms'), json.dumps(r.get('small_chi_variant'))[:900])
PY
