# Round-2 profile recipe (B200_PROFILING.md): full captures of the block search and the pair sweeps.
# usage: bash scripts/profile_r2.sh <particles> <tag> [kernels...]   (reports are reduced to CSV pages on the box: the .ncu-rep files are big)
N=${1:-8e6}; TAG=${2:-8M}; shift 2
mkdir -p gpurun_out
prof() {  # name, kernel regex
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${SKIP:-4} -c 1 -f -o /tmp/prof_$1 \
      python bench.py --particles $N --steps 1 --warmup 3 --no-e2e --no-cpu > /tmp/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_${TAG}_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2_${TAG}_$1_source.csv 2>/dev/null
  tail -1 /tmp/ncu_$1.log | cut -c1-200
}
for K in "$@"; do
  case $K in
    search) prof search 'k_block_search' ;;
    density) SKIP=12 prof density k_sweep ;;
    accv) SKIP=13 prof accv k_sweep ;;
    dedt) SKIP=14 prof dedt k_sweep ;;
  esac
done
ls -la gpurun_out | grep r2_${TAG}
