mkdir -p gpurun_out
for v in refine norefine; do if [ $v = norefine ]; then export T4B_SVD_NOREFINE=1; fi
 timeout 300 python bench.py --workload c5 --steps 2 --warmup 1 > gpurun_out/bench_r02y_c5_$v.json 2> gpurun_out/bench_r02y_c5_$v.err
 timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 > gpurun_out/bench_r02y_c2_$v.json 2> gpurun_out/bench_r02y_c2_$v.err
done
python - <<'PY'
import json
for v in ('refine','norefine'):
    for w in ('c5','c2'):
        try:
            d=json.loads(open('gpurun_out/bench_r02y_%s_%s.json'%(w,v)).read().strip().splitlines()[-1]); r=d.get('record',d)
            if w=='c5': print(w,v, r.get('value'), r.get('ms'), r.get('phase_ms'), json.dumps(r.get('small_chi_variant'))[:600])
            else: print(w,v, r.get('ms_per_apply'), {k:round(x,1) for k,x in r['kernel_profile_ms'].items() if x>8})
        except Exception as e: print('ERR',w,v,e)
PY
