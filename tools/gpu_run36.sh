mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rrlu.py tests/test_gpu_linalg_extra.py tests/test_gpu_tci2.py tests/test_gpu_treetci.py tests/test_gpu_simplett.py -m gpu -q -x 2>&1 | grep -v "^\[t4b\]" | tail -6
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['record']: print(r['shape'], round(r['ms'],2), r['roofline']['achieved'], r['kernel_profile_ms'])
"
