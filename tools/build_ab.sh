#!/bin/bash
# A/B build of the library: the product sources plus the FFMA engine of the step kernel (csrc/pinn_step_ffma.cu) and the
# environment switches PINN_B200_ENGINE / PINN_B200_HOST_ZEROCOPY / PINN_B200_HOST_INLINE (-DPINN_AB_BUILD).  Not loaded
# by the package unless PINN_B200_LIBRARY points at it:
#   tools/build_ab.sh && PINN_B200_LIBRARY=tools/dbg/libpinn_b200_ab.so PINN_B200_ENGINE=ffma python bench.py
set -e
cd "$(dirname "$0")/../pinn_for_quantum_wavefunction_surfaces_b200/csrc"
mkdir -p ../../tools/dbg
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -DPINN_AB_BUILD "$@" \
     -o ../../tools/dbg/libpinn_b200_ab.so pinn_reduce.cu pinn_step_ffma.cu pinn_step_tc.cu pinn_train.cu pinn_capi.cu \
     pinn_train_api.cu -lcudart
echo built tools/dbg/libpinn_b200_ab.so
