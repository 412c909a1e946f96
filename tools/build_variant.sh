#!/bin/bash
# Builds an A/B variant of the library: tools/build_variant.sh <name> <nvcc -D flags...>
#   -> genedex_b200/csrc/variants/lib<name>.so   (select it with GENEDEX_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/../genedex_b200/csrc"
name=$1; shift
mkdir -p variants
make -s device_build.o host_build.o host_pack.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function \
    --expt-relaxed-constexpr "$@" -c api.cu -o variants/api_$name.o 2> variants/$name.log
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib$name.so variants/api_$name.o device_build.o host_build.o host_pack.o \
    -lcudart_static -lpthread -ldl -lrt
rm -f variants/api_$name.o
echo "built genedex_b200/csrc/variants/lib$name.so"
