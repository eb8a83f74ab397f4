#!/bin/bash
# scripts/build_variant.sh NAME "EXTRA NVCC FLAGS"  ->  mvs_gi_b200/libmvsgi_b200_NAME.so (std builder, 3/4 cameras only)
set -e
cd "$(dirname "$0")/../mvs_gi_b200/csrc"
NAME=$1; EXTRA=$2
OBJ=../../build/csrc/var_$NAME; mkdir -p $OBJ
FL="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DMVSGI_BENCH_ONLY $EXTRA"
nvcc $FL -c -o $OBJ/mvsgi_b200.o mvsgi_b200.cu &
nvcc $FL -c -o $OBJ/sweep_cta.o sweep_cta.cu &
nvcc $FL -Xptxas -v -c -o $OBJ/sweep_cta_fb.o sweep_cta_fb.cu > $OBJ/ptxas.log 2>&1 &
nvcc $FL -c -o $OBJ/softargmin_tma.o softargmin_tma.cu &
nvcc $FL -c -o $OBJ/softargmin_bwd.o softargmin_bwd.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libmvsgi_b200_$NAME.so $OBJ/mvsgi_b200.o $OBJ/sweep_cta.o $OBJ/sweep_cta_fb.o $OBJ/softargmin_tma.o $OBJ/softargmin_bwd.o
grep -A2 "sweep_kernelILi3ELi0ELb1ELb0ELb0ELb1ELb0ELb1ELb0ELi[0-9]E" $OBJ/ptxas.log | grep "registers" | tr '\n' ' '; echo " <- $NAME"
