#!/bin/bash
# usage: variant_bench.sh "<nvcc -D flags>" [bench args]: rebuild the library with the flags on the GPU box is impossible
# (no nvcc need: nvcc exists there too) -> build here, run there.
set -e
