#!/bin/bash
# usage: tools/ladder_census.sh [tag] [extra nvcc flags]  -> compiles the stand-alone ladder TU (same as order_search.py) against
# the CURRENT headers and prints the opcode census of its hot loop (tools/sass_census.py)
set -e
cd "$(dirname "$0")/.."
tag=${1:-cur}; shift || true
mkdir -p build/scratch
python3 - "$tag" <<'PY'
import sys, os
sys.path.insert(0, "tools")
import order_search
open("build/scratch/ladder_tu_%s.cu" % sys.argv[1], "w").write(order_search.TU)
PY
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -diag-suppress 550 -cubin "$@" -o build/scratch/ladder_tu_$tag.cubin build/scratch/ladder_tu_$tag.cu
cuobjdump -sass build/scratch/ladder_tu_$tag.cubin > build/scratch/ladder_tu_$tag.sass
python3 tools/sass_census.py build/scratch/ladder_tu_$tag.sass k_ladder
