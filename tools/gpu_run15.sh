mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^\[t4b\]" | tail -80 > gpurun_out/pytest_r02o.log
grep -E "C3 saturated|C3 full sweep|C2 full|passed|failed|FAILED" gpurun_out/pytest_r02o.log | tail -12
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02o.json 2> gpurun_out/bench_r02o.err
T4B_VERBOSE=1 timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5 2> gpurun_out/verbose_r02o.err > /dev/null
grep -E "Cholesky|kappa" gpurun_out/verbose_r02o.err | sed -E 's/[0-9]\.[0-9]+e[-+][0-9]+/X/g' | sort | uniq -c | sort -rn | head -14
grep -E "kappa" gpurun_out/verbose_r02o.err | sed -E 's/.*kappa ([0-9.e+-]+) .*/\1/' | sort -g | awk '{a[NR]=$1} END {print "kappa min/median/max", a[1], a[int(NR/2)+1], a[NR], NR}'
grep -E "defect" gpurun_out/verbose_r02o.err | sed -E 's/.*defect ([0-9.e+-]+).*/\1/' | sort -g | awk '{a[NR]=$1} END {print "defect min/median/max", a[1], a[int(NR/2)+1], a[NR], NR}'
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 > gpurun_out/bench_r02o_c4.json 2> gpurun_out/bench_r02o_c4.err
python - <<'PY'
import json
def rec(f):
    return json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
try:
    d=rec('bench_r02o'); print('c3', round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], d['kernel_profile_ms'])
except Exception as e: print('c3 ERR',e)
try:
    d=rec('bench_r02o_c4')['record']; print('c4', [(r['shape'], round(r['ms'],2), round(r['roofline']['frac'],3)) for r in d])
except Exception as e: print('c4 ERR',e)
PY
