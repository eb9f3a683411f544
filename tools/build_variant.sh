#!/bin/bash
# A/B builds: tools/build_variant.sh NAME "-DFOO=1 ..." -> recoup_b200/variants/lib_NAME.so
# (select with RCP_LIB_PATH=recoup_b200/variants/lib_NAME.so; the directory is git-ignored)
set -e
cd "$(dirname "$0")/../recoup_b200/csrc"
name=$1; flags=$2
bd=build_var/$name
mkdir -p $bd ../variants
for f in api bgzf import index sort scan coverage coverage_buckets coverage_split profile consumers; do
  if [ "$f" = coverage_split ] || [ "$f" = profile ] || [ ! -f build/$f.o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. $flags -c $f.cu -o $bd/$f.o &
  else
    cp build/$f.o $bd/$f.o
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$name.so $bd/*.o -lcudart -lz
echo built ../variants/lib_$name.so
