#!/bin/bash
# usage: tools/build_variant.sh <tag> "<extra nvcc flags>"  ->  cudaraytracing_b200/variants/libcrt_<tag>.so (select with CRT_LIB)
set -e
tag=$1; flags=$2
cd "$(dirname "$0")/../cudaraytracing_b200"
mkdir -p variants _build/var_$tag
make -s OBJDIR=_build/var_$tag EXTRA_NVFLAGS="$flags" _build/var_$tag/crt_bvh_build.o _build/var_$tag/crt_render.o _build/var_$tag/crt_api.o \
     _build/var_$tag/crt_host.o _build/var_$tag/crt_config_png.o 2>&1 | grep -E "error|ptxas info" || true
nvcc -shared -o variants/libcrt_$tag.so _build/var_$tag/*.o -lz -cudart shared
echo built variants/libcrt_$tag.so
