set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 57 -c 1 -o gpurun_out/prof_dedt \
    python bench.py --particles 8e6 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu3.log 2>&1
