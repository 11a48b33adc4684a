mkdir -p gpurun_out
T4B_VERBOSE=2 timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-c5 2> gpurun_out/verbose2_r02t.err > /dev/null
grep -B1 "nx=512 npad=512" gpurun_out/verbose2_r02t.err | grep "off per sweep" | awk '{print $5}' | sort -g | awk '{a[NR]=$1} END {print "n=512 first-sweep off: min/median/max", a[1], a[int(NR/2)+1], a[NR], NR}'
grep -B1 "nx=512 npad=512" gpurun_out/verbose2_r02t.err | grep "off per sweep" | head -5
grep -B1 "nx=2048 npad=2048" gpurun_out/verbose2_r02t.err | grep "off per sweep" | head -3
