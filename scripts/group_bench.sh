#!/bin/bash
# Builds bench_cpp/group_search.cpp against the in-tree libvsm.so and runs it: scripts/group_bench.sh <n_gpus> [rows] [nq] [steps]
set -e
cd "$(dirname "$0")/.."
LIB=$PWD/visual-slam-pipeline_b200/lib
mkdir -p /tmp/vsm_bench
g++ -O2 -std=c++17 -I include -I /usr/local/cuda/include bench_cpp/group_search.cpp -L "$LIB" -lvsm \
    -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$LIB" -lpthread -ldl -o /tmp/vsm_bench/group_search
/tmp/vsm_bench/group_search "$@"
