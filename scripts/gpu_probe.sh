# box facts + the library's micro-benchmarks (fp64 DFMA peak, copy bandwidth)
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; } > gpurun_out/box_info.txt 2>&1
python - <<'PY' > gpurun_out/microbench.txt 2>&1
from flecsph_b200 import capi
lib = capi.load(capi.PRODUCT_LIB)
print("fp64_tflops", lib.fsph_measure_fp64_tflops(0, 5))
print("copy_gbs", lib.fsph_measure_copy_gbs(0, 4 << 30, 5))
PY
cat gpurun_out/box_info.txt gpurun_out/microbench.txt
