#!/bin/bash
# build_variant.sh <name> <extra nvcc flags...>  ->  build/variants/libtdl_<name>.so   (kernel experiments)
set -e
PKG=$(ls -d /root/repo/tripled*_b200); OUT=/root/repo/build/variants; mkdir -p $OUT/$1
name=$1; shift
for f in tdl_api tdl_photo tdl_smooth tdl_feat; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I /root/repo/include -I $PKG/csrc "$@" -c $PKG/csrc/$f.cu -o $OUT/$name/$f.o &
done; wait
nvcc -shared -o $OUT/libtdl_$name.so $OUT/$name/*.o -gencode arch=compute_100a,code=sm_100a
echo built $OUT/libtdl_$name.so
