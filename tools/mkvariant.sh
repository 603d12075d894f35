#!/bin/bash
# usage: tools/mkvariant.sh <name> [extra nvcc flags...]  -> build/variants/<name>.so (+ ptxas line of the C2 kernel)
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC -shared -ldl -I include -Xptxas -v "$@" -o build/variants/$name.so occlusionenv_b200/csrc/occl_b200.cu 2>&1 \
  | awk '/Compiling entry function/ {name=$0} /Used [0-9]+ registers/ {if (name ~ /raster_kernelILb0ELi32ELi32ELb0/ || name ~ /raster_kernelILb0ELi128ELi4ELb0/ || name ~ /raster_kernelILb1ELi32ELi32ELb0/) print substr(name, index(name,"_Z"), 40), $0; name=""} /spill/ {sp=$0}'
