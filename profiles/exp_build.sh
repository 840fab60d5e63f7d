#!/bin/bash
# Experiment build (CPU container): one kernel instantiation (cfg3 fp16 shape) plus extra -D switches.
#   profiles/exp_build.sh <name> [-DFOO ...]  ->  mdy-newanalysis-package_b200/newanalysis/exp/libgfn_<name>.so
set -e
cd "$(dirname "$0")/../mdy-newanalysis-package_b200/csrc"
name=$1; shift
mkdir -p ../newanalysis/exp
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -warn-spills"
$NV ${ONLY:--DGFN_EXP_ONLY_H16} "$@" -c -o /tmp/exp_$name.o gfn_kernels.cu
make -s gfn_api.o gfn_traj.o gfn_dense.o gfn_count.o
API=gfn_api.o
case "${ONLY:-} $*" in *GFN_EXP_12_6*)      # the host side pads selections with the same constants (gfn_kernels.cuh)
  $NV -DGFN_EXP_12_6 -c -o /tmp/exp_api_$name.o gfn_api.cu; API=/tmp/exp_api_$name.o;;
esac
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../newanalysis/exp/libgfn_$name.so /tmp/exp_$name.o $API gfn_traj.o gfn_dense.o gfn_count.o -cudart static -ldl -lpthread
echo built $name
