mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^\[t4b\]" | tail -80 > gpurun_out/pytest_r02p.log
grep -E "C3 saturated|C3 full sweep|C2 full|passed|failed|FAILED" gpurun_out/pytest_r02p.log | tail -12
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 > gpurun_out/bench_r02p.json 2> gpurun_out/bench_r02p.err
for b in 0 1 2 3 4; do T4B_RRLU_BPS=$b timeout 300 python bench.py --workload c4 --steps 5 --warmup 2 > gpurun_out/bench_r02p_c4_bps$b.json 2> gpurun_out/bench_r02p_c4_bps$b.err; done
timeout 600 python bench.py --workload c2 --steps 2 --warmup 1 > gpurun_out/bench_r02p_c2.json 2> gpurun_out/bench_r02p_c2.err
python - <<'PY'
import json
def rec(f):
    return json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
try:
    d=rec('bench_r02p'); print('c3', round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], d['kernel_profile_ms'])
except Exception as e: print('c3 ERR',e)
for b in range(5):
    try:
        d=rec('bench_r02p_c4_bps%d'%b)['record']; print('c4 bps',b, [(r['shape'], round(r['ms'],2)) for r in d])
    except Exception as e: print('c4 ERR',e)
try:
    d=rec('bench_r02p_c2')['record']; print('c2', d['ms_per_apply'], d['kernel_profile_ms'])
except Exception as e: print('c2 ERR',e)
PY
