#!/bin/bash
# developer build: product library + host simulation + profiling / experiment variants under build/ (git-ignored)
# usage: tools/_dev/build_all.sh [variant flags...]   e.g.  tools/_dev/build_all.sh prof1:-DGUSTO_PROF_MODE=1 nodmma:-DGUSTO_NO_DMMA
set -e
R=/root/repo
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared"
mkdir -p $R/build $R/tests/hostsim/_build
g++ -O2 -std=c++17 -fPIC -shared -x c++ $R/tests/hostsim/hostsim.cpp -o $R/tests/hostsim/_build/libgusto_hostsim.so
$NV -Xptxas -v -o $R/gusto.jl_b200/libgusto_b200.so $R/gusto.jl_b200/csrc/capi.cu > $R/build/ptxas.log 2>&1 &
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  $NV ${flags//,/ } -o $R/build/libgusto_$name.so $R/gusto.jl_b200/csrc/capi.cu > $R/build/ptxas_$name.log 2>&1 &
done
wait
grep -A2 "ipm_kernelILi2E\|ipm_kernelILi3E" $R/build/ptxas.log | grep "stack\|registers"
