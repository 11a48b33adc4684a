mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^\[t4b\]" | tail -80 > gpurun_out/pytest_r02n.log
grep -E "C3 saturated|C3 full sweep|C2 full|passed|failed|FAILED" gpurun_out/pytest_r02n.log | tail -12
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02n.json 2> gpurun_out/bench_r02n.err
timeout 600 python bench.py --workload c2 --steps 2 --warmup 1 > gpurun_out/bench_r02n_c2.json 2> gpurun_out/bench_r02n_c2.err
T4B_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5 2>&1 | grep -E "Cholesky|kappa" | sed -E 's/[0-9]\.[0-9]+e[-+][0-9]+/X/g' | sort | uniq -c | sort -rn | head -12
T4B_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5 2>&1 | grep -E "kappa" | awk '{print $8}' | sort -g | awk '{a[NR]=$1} END {print "kappa min/median/max", a[1], a[int(NR/2)], a[NR], NR}'
python - <<'PY'
import json
def rec(f):
    return json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
try:
    d=rec('bench_r02n'); print('c3', round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], d['kernel_profile_ms'])
except Exception as e: print('c3 ERR',e)
try:
    d=rec('bench_r02n_c2')['record']; print('c2', d['ms_per_apply'], d['kernel_profile_ms'])
except Exception as e: print('c2 ERR',e)
PY
