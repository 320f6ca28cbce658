#!/bin/bash
# Builds libugrt.so (sm_100a only, in-tree next to the package).  -fmad=false: parity with the
# reference's arithmetic forbids FMA contraction.
set -e
cd "$(dirname "$0")"
OUT=${UG_OUT:-../libugrt.so}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-Wall -I../../include -I. ${UG_EXTRA_FLAGS}"
SRCS="context.cu k_sort.cu k_grid.cu k_gridfused.cu k_gridsweep.cu k_trace.cu k_shadow.cu k_shadowdirect.cu k_lightshare.cu k_shade.cu k_peak.cu k_reflect.cu k_ugridfused.cu camera_host.cpp obj_loader.cpp ref_compat.cpp comm.cpp present_gl.cpp"
$NVCC $FLAGS ${UG_PTXAS_V:+-Xptxas -v} -shared -o $OUT $SRCS -lcudart -ldl
echo "built $(readlink -f $OUT)"
