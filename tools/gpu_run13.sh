mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^\[t4b\]" | tail -60 > gpurun_out/pytest_r02m.log
grep -E "C3 saturated|C3 full sweep|C2 full|passed|failed|FAILED" gpurun_out/pytest_r02m.log | tail -12
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02m.json 2> gpurun_out/bench_r02m.err
timeout 600 python bench.py --workload c2 --steps 2 --warmup 1 > gpurun_out/bench_r02m_c2.json 2> gpurun_out/bench_r02m_c2.err
T4B_SVD_NOGRAM=1 timeout 600 python bench.py --workload c2 --steps 2 --warmup 1 > gpurun_out/bench_r02m_c2_nogram.json 2> gpurun_out/bench_r02m_c2_nogram.err
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 > gpurun_out/bench_r02m_c4.json 2> gpurun_out/bench_r02m_c4.err
T4B_RRLU_BPS=1 timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 > gpurun_out/bench_r02m_c4_bps1.json 2> gpurun_out/bench_r02m_c4_bps1.err
T4B_RRLU_BPS=2 timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 > gpurun_out/bench_r02m_c4_bps2.json 2> gpurun_out/bench_r02m_c4_bps2.err
timeout 300 python bench.py --workload c1 --steps 3 --warmup 2 > gpurun_out/bench_r02m_c1.json 2> gpurun_out/bench_r02m_c1.err
for w in 1 2 8; do T4B_PATCH_WORKERS=$w timeout 300 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_r02m_c5_w$w.json 2> gpurun_out/bench_r02m_c5_w$w.err; done
python - <<'PY'
import json
def rec(f):
    return json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
try:
    d=rec('bench_r02m'); print('c3', round(d['ms_per_step'],1), d['e2e']['ms_per_step'], d['roofline_contraction']['frac'], d['kernel_profile_ms'], 'c5', d.get('c5',{}).get('value'))
except Exception as e: print('c3 ERR',e)
for f in ['c2','c2_nogram']:
    try:
        d=rec('bench_r02m_'+f)['record']; print(f, d['ms_per_apply'], d['kernel_profile_ms'])
    except Exception as e: print(f,'ERR',e)
for f in ['c4','c4_bps1','c4_bps2']:
    try:
        d=rec('bench_r02m_'+f)['record']; print(f, [(r['shape'], round(r['ms'],2)) for r in d])
    except Exception as e: print(f,'ERR',e)
try:
    d=rec('bench_r02m_c1')['record']; print('c1', d['batch1']['ms_per_compress'], d['batched']['ms_per_batch'])
except Exception as e: print('c1 ERR',e)
for w in [1,2,8]:
    try:
        d=rec('bench_r02m_c5_w%d'%w)['record']; print('c5 workers',w, d['value'], d['phase_ms'])
    except Exception as e: print('c5',w,'ERR',e)
PY
