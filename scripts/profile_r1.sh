# ncu launch list + full capture of the dominant kernels (1 GPU, small step count)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --particles 8e6 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_sweep|k_search' -s 100 -c 24 -o gpurun_out/prof_r1 \
    python bench.py --particles 8e6 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
