#!/bin/bash
# A/B builds of the lean kernel: scripts/variants.sh NAME "-DFLAG ..." -> hnswlib-rs_b200/lib/variants/libhnsw_b200_NAME.so
# (only the OpL2 / 512-byte-row / ef<=64 instantiation); run with HNSW_B200_LIB=<that file> python scripts/quick_search.py
set -e
cd "$(dirname "$0")/../hnswlib-rs_b200/csrc"
name=$1; flags=$2
mkdir -p ../lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -DHB_FAST_BUILD $flags -c search_lean_f32.cu -o ../lib/variants/search_lean_$name.o 2> ../lib/variants/$name.ptxas.log
grep -A3 "OpL2ELi4ELi64ELb0" ../lib/variants/$name.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
objs=$(ls ../lib/obj/*.o | grep -v search_lean_f32.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libhnsw_b200_$name.so $objs ../lib/variants/search_lean_$name.o -lcudart_static -lpthread -ldl -lrt
