# Round-2 profile recipe (B200_PROFILING.md).  usage: bash scripts/profile_r2.sh <particles> <tag> [launches] [traffic] [search] [walk] [density] [accv] [dedt]
#   launches : every launch of the last step with its device time (gpu__time_duration.sum)
#   traffic  : DRAM bytes of the walk, the pair kernel of the search and the three pair sweeps of the last step (one cheap pass)
#   search / density / accv / dedt : ncu --set full of that kernel in the last step; the .ncu-rep files are big, so the
#              raw and source pages are exported to CSV on the box (gpurun_out/r2_<tag>_<kernel>_{raw,source}.csv)
N=${1:-8e6}; TAG=${2:-8M}; shift 2
BENCH="python bench.py --particles $N --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs"
mkdir -p gpurun_out
prof() {  # name, kernel base-name regex, launches to skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o /tmp/prof_$1 $BENCH > /tmp/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_${TAG}_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2_${TAG}_$1_source.csv 2>/dev/null
  tail -1 /tmp/ncu_$1.log | cut -c1-160
}
# per step: 1 k_block_walk, 1 k_block_pairs, 3 k_sweep (density, acceleration + signal speed, de/dt); steps run: initial + 3 warm-up + 1
for K in "$@"; do
  case $K in
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_${TAG}_launches.csv $BENCH > /tmp/ncu_launches.log 2>&1 ;;
    traffic) ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:k_block_walk|k_block_pairs|k_sweep" -s 20 -c 5 --csv \
               --log-file gpurun_out/r2_${TAG}_traffic.csv $BENCH > /tmp/ncu_traffic.log 2>&1 ;;
    search) prof search k_block_pairs 4 ;;
    walk) prof walk k_block_walk 4 ;;
    density) prof density k_sweep 12 ;;
    accv) prof accv k_sweep 13 ;;
    dedt) prof dedt k_sweep 14 ;;
  esac
done
ls -la gpurun_out | grep r2_${TAG}
