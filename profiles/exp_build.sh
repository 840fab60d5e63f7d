#!/bin/bash
# Experiment build (CPU container): one pair_kernel instantiation (cfg3 fp16 shape) plus extra -D switches.
#   profiles/exp_build.sh <name> [-DFOO ...]  ->  mdy-newanalysis-package_b200/newanalysis/exp/<name>/libgfn.so
# On the GPU box the shim picks it up through LD_LIBRARY_PATH (it is searched before the shim's RUNPATH $ORIGIN):
#   LD_LIBRARY_PATH=$PWD/mdy-newanalysis-package_b200/newanalysis/exp/<name> timeout 60 python bench.py --no-other-configs --no-cpu-baseline
# ONLY=-DGFN_EXP_12_6 builds the 12 filter + 6 drain warp shape (768-site core tiles, 1,024-entry rings; the host side is
# compiled with the same switch because selections are padded to 3,072 sites).  Measured in round 2 (cfg3, B200):
#   -DGFN_REGS_F6='"72"' -DGFN_REGS_D6='"128"'   1,349 x10^9 pair evals/s (0.48 of the FP64 peak) against 1,560 for 8 + 8 - parity exact
#   -DGFN_REGS_F6='"80"' -DGFN_REGS_D6='"128"'   hangs: 384*80 + 192*128 is exactly the 576 x 96 registers the CTA launches with, and
#                                               the drain warps' setmaxnreg.inc never gets its registers
set -e
cd "$(dirname "$0")/../mdy-newanalysis-package_b200/csrc"
name=$1; shift
mkdir -p ../newanalysis/exp/$name
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -warn-spills"
$NV ${ONLY:--DGFN_EXP_ONLY_H16} "$@" -c -o /tmp/exp_$name.o gfn_kernels.cu
make -s gfn_api.o gfn_traj.o gfn_dense.o gfn_count.o gfn_cells.o
API=gfn_api.o
case "${ONLY:-} $*" in *GFN_EXP_12_6*)      # the host side pads selections with the same constants (gfn_kernels.cuh)
  $NV -DGFN_EXP_12_6 -c -o /tmp/exp_api_$name.o gfn_api.cu; API=/tmp/exp_api_$name.o;;
esac
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../newanalysis/exp/$name/libgfn.so /tmp/exp_$name.o $API gfn_traj.o gfn_dense.o gfn_count.o gfn_cells.o -cudart static -ldl -lpthread
echo built $name
