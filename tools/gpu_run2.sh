mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -40 > gpurun_out/pytest_r02b.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err
for wl in c4 c2 c1; do timeout 600 python bench.py --workload $wl --steps 3 --warmup 2 > gpurun_out/bench_r02b_$wl.json 2> gpurun_out/bench_r02b_$wl.err; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02b_ref.json 2> gpurun_out/bench_r02b_ref.err
tail -15 gpurun_out/pytest_r02b.log; tail -c 600 gpurun_out/bench_r02b.err; nproc
