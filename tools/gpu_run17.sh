mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_patches_nccl.py tests/test_partitioned_contract.py tests/test_gpu_patches.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_r02q_nccl.log
tail -5 gpurun_out/pytest_r02q_nccl.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02q_2gpu.json 2> gpurun_out/bench_r02q_2gpu.err
tail -c 400 gpurun_out/bench_r02q_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r02q_2gpu.json').read().strip().splitlines()[-1])
    print('c3 x2', d['n_gpus'], round(d['ms_per_step'],1), d['value'], d['e2e'])
    print('c5', d.get('c5'))
except Exception as e: print('ERR', e)
PY
