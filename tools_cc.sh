#!/bin/bash
# usage: tools_cc.sh <file.cu> [extra nvcc flags]: compile one translation unit, print registers / spills per kernel
f=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -I ../../include -I . "$@" -c $f -o /tmp/$(basename $f .cu).o 2>&1 | grep -E "error|warning|Compiling entry|spill|Used" | sed -E 's/.*Compiling entry function .([A-Za-z0-9_]+). for.*/\1/' | c++filt | cut -c1-150
