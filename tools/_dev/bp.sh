#!/bin/bash
# build the library, then probe the solve kernel on a B200 (developer helper)
set -e
cd /root/repo
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v $NVCC_EXTRA -o gusto.jl_b200/libgusto_b200.so gusto.jl_b200/csrc/capi.cu 2>&1 | grep -A3 "Compiling entry function '_Z10ipm_kernelILi2" | tail -2
/usr/local/graft/bin/gpurun --timeout 300 -- "python tools/gpu_probe.py ${1:-astrobeeSE3} ${2:-1024} 2 ${3}" 2>&1 | grep -v "^\[gpurun\] sending"
