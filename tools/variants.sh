#!/bin/bash
# Build variants of libxsq.so that differ in the XSQ_V_* macros of xsq_rk_fast.cuh
# (only the Ts5 / CK5 units are recompiled): extensisq_b200/libxsq_<name>.so.
# usage: tools/variants.sh name "-DXSQ_V_PREFETCH=1 ..."
set -e
cd "$(dirname "$0")/../extensisq_b200/csrc"
name=$1; flags=$2
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off -I. -I../../include"
mkdir -p build/var_$name
for t in Ts5 CK5; do $NV $flags -DXSQ_INST_TAB=$t -c xsq_rk_inst.cu -o build/var_$name/inst_$t.o & done
wait
objs=$(ls build/*.o | grep -v "inst_Ts5.o\|inst_CK5.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libxsq_$name.so $objs build/var_$name/inst_Ts5.o build/var_$name/inst_CK5.o -ldl
ls -la ../libxsq_$name.so
