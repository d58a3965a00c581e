#!/bin/bash
# usage: tools/probe.sh file.cu [extra nvcc flags]   -> compiles to cubin, prints loop stats and the toy issue model
set -e
cd /root/repo
src=$1; shift
out=build/scratch/$(basename $src .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -diag-suppress 550 -cubin -o $out.cubin $src "$@"
cuobjdump -sass $out.cubin > $out.sass
cd tools
for k in $(grep -o "Function : [A-Za-z0-9_]*" ../$out.sass | awk '{print $3}'); do
  line=$(python3 sass_stalls.py ../$out.sass $k 2>/dev/null | head -1)
  echo "$k: $line"
  python3 sass_sim.py ../$out.sass $k $(echo "$line" | sed -E 's/loop 0x([0-9a-f]+)\.\.0x([0-9a-f]+).*/\1 \2/') 4 gto
done
