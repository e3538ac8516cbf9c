#!/bin/bash
# Instrumented build of the library for tools/conv_trace.py: the same sources with -DSMB_TRACE (per-role clock64 stamps of
# CTA 0 in conv_gemm_kernel).  Output: tools/_trace/libsipmask_b200_trace.so (git-ignored, travels with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_trace/obj
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
for f in sipmask_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  if [ "$b" = conv_tcgen05 ]; then
    $NV -DSMB_TRACE -c $f -o tools/_trace/obj/$b.o &
  else
    cp sipmask_b200/lib/obj/$b.o tools/_trace/obj/$b.o
  fi
done
wait
/usr/local/cuda/bin/nvcc -shared -o tools/_trace/libsipmask_b200_trace.so tools/_trace/obj/*.o -gencode arch=compute_100a,code=sm_100a
ls -la tools/_trace/libsipmask_b200_trace.so
