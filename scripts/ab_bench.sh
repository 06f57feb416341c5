#!/bin/bash
# A/B of kernel variants on ONE box: builds variants/lib_<name>.so here (nvcc cross-compiles), then the gpurun command
# runs bench.py once per library.  usage: scripts/ab_bench.sh "<bench args>" name1 "<defs1>" name2 "<defs2>" ...
set -e
cd "$(dirname "$0")/.."
ARGS="$1"; shift
mkdir -p variants
CMD=""
while [ $# -gt 0 ]; do
  NAME="$1"; DEFS="$2"; shift 2
  FSPH_NVCC_DEFS="$DEFS" FSPH_BUILD_OUT="$PWD/variants/lib_$NAME.so" FSPH_OBJ_SUFFIX="_$NAME" python flecsph_b200/build.py > /dev/null
  CMD="$CMD echo == $NAME: $DEFS; FSPH_PRODUCT_LIB=variants/lib_$NAME.so python bench.py $ARGS 2>&1 | python scripts/show_bench.py;"
done
/usr/local/graft/bin/gpurun --timeout 900 -- "$CMD"
