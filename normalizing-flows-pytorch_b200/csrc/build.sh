#!/bin/bash
# Build libnfb200.so for sm_100a (B200).  Usage: ./build.sh [EXTRA="extra nvcc flags"]
set -e
cd "$(dirname "$0")"
make -j"$(nproc)" "$@"
echo "built $(cd .. && pwd)/libnfb200.so"
