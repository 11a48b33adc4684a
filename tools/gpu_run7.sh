mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_svd.py tests/test_gpu_simplett.py tests/test_gpu_seam.py tests/test_gpu_patches.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_r02g_svd.log
tail -4 gpurun_out/pytest_r02g_svd.log
for v in default nobatch lpp4 lpp8 lpp16 lpp32; do
  case $v in
    default) E="";;
    nobatch) E="T4B_SVD_NOBATCH=1";;
    lpp4) E="T4B_SVD_LPP=4";;
    lpp8) E="T4B_SVD_LPP=8";;
    lpp16) E="T4B_SVD_LPP=16";;
    lpp32) E="T4B_SVD_LPP=32";;
  esac
  env $E timeout 300 python tools/probe_svd_small.py > gpurun_out/probe_svd_small_$v.json 2> gpurun_out/probe_svd_small_$v.err
done
timeout 300 python bench.py --workload c1 --steps 3 --warmup 2 > gpurun_out/bench_r02g_c1.json 2> gpurun_out/bench_r02g_c1.err
timeout 300 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_r02g_c5.json 2> gpurun_out/bench_r02g_c5.err
T4B_SVD_NOBATCH=1 timeout 300 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_r02g_c5_nobatch.json 2> gpurun_out/bench_r02g_c5_nobatch.err
python - <<'PY'
import json
vs=['default','nobatch','lpp4','lpp8','lpp16','lpp32']
d={}
for v in vs:
    try: d[v]=json.load(open('gpurun_out/probe_svd_small_%s.json'%v))
    except Exception as e: print(v,'ERR',e)
keys=list(next(iter(d.values())).keys()) if d else []
print('%-12s'%'shape'+''.join('%10s'%v for v in d))
for k in keys:
    print('%-12s'%k+''.join('%10s'%d[v].get(k,{}).get('us') for v in d), ' err', max(d[v][k]['sigma_err'] for v in d if k in d[v]))
for f in ['c1','c5','c5_nobatch']:
    try: print(open('gpurun_out/bench_r02g_%s.json'%f).read()[:1500])
    except Exception as e: print(f,'ERR',e)
PY
