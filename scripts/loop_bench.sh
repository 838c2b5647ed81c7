#!/bin/bash
# Builds bench_cpp/loop_search.cpp against the in-tree libvsm.so and runs it: scripts/loop_bench.sh [nkf rows nq every steps]
set -e
cd "$(dirname "$0")/.."
LIB=$PWD/visual-slam-pipeline_b200/lib
mkdir -p /tmp/vsm_bench
g++ -O2 -std=c++17 -I include -I /usr/local/cuda/include bench_cpp/loop_search.cpp -L "$LIB" -lvsm \
    -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$LIB" -lpthread -ldl -o /tmp/vsm_bench/loop_search
/tmp/vsm_bench/loop_search "$@"
