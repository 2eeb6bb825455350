#!/bin/bash
# Builds the product library again with extra -D flags into aws-c-compression_b200/lib/variants/<name>.so
# (A/B runs on the GPU box: AWS_HUFFMAN_B200_LIB=<that file> python bench.py ...).
#   tools/build_variant.sh <name> "-DHB_DEC_UNIFIED=0 ..."
set -e
name=$1; defs=$2
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=$root/aws-c-compression_b200
mkdir -p $pkg/lib/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $defs \
    -I $root/include -I $root/shim/aws-c-common/include -c $pkg/csrc/huffman_batch.cu -o /tmp/hb_variant_$name.o
objs=$(ls $pkg/build/*.c.o)
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -Xlinker -Bsymbolic -o $pkg/lib/variants/$name.so $objs /tmp/hb_variant_$name.o
echo built $pkg/lib/variants/$name.so
