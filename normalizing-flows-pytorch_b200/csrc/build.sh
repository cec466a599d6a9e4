#!/bin/bash
# Build libnfb200.so for sm_100a (B200).  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=../libnfb200.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -shared -I../../include \
     --threads 4 "$@" \
     -o "$OUT" *.cu
echo "built $(cd .. && pwd)/libnfb200.so"
