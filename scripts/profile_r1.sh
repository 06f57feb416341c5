# Round-1 profile recipe (B200_PROFILING.md): launch list of the bench command + full captures of the two top kernels.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_r1.log 2>&1
# dedt sweep, size class 2 (the most populated class at 64 M), and the neighbour search of the same class, at 8 M
# particles (same tree structure as 64 M: one level shallower), one launch each from the last step
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 57 -c 1 -o gpurun_out/prof_r1_dedt \
    python bench.py --particles 8e6 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_search -s 21 -c 1 -o gpurun_out/prof_r1_search \
    python bench.py --particles 8e6 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ls -la gpurun_out
