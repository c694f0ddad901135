#!/bin/bash
# Build a tuning variant of the product library next to the default one: tools/build_variant.sh <tag> <extra nvcc flags...>
# -> gpurun_variants/libportello_b200_<tag>.so   (load it with PORTELLO_B200_LIB=...)
set -e
TAG=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/portello_b200/csrc
mkdir -p $ROOT/gpurun_variants/_obj_$TAG
O=$ROOT/gpurun_variants/_obj_$TAG
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
nvcc $F "$@" -c $C/device/kernels.cu -o $O/kernels.o
nvcc $F "$@" -c $C/context.cu -o $O/context.o
g++ -O3 -std=c++17 -fPIC -pthread -c $C/host/pack.cpp -o $O/pack.o
g++ -O3 -std=c++17 -fPIC -pthread -c $C/host/contig_prep.cpp -o $O/contig_prep.o
for f in bgzf bam_io bam_abi; do g++ -O3 -std=c++17 -fPIC -pthread -c $C/host/$f.cpp -o $O/$f.o; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/gpurun_variants/libportello_b200_$TAG.so $O/kernels.o $O/context.o $O/pack.o $O/contig_prep.o $O/bgzf.o $O/bam_io.o $O/bam_abi.o -cudart static -lz
rm -rf $O
