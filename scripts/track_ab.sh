#!/bin/bash
# Tracking-step latency from C++ (bench_cpp/track_latency.cpp) with the tile top-2 records (default) and with the
# top-4 records (VSM_NO_T2=1).  Usage: scripts/track_ab.sh [pairs]
set -e
cd "$(dirname "$0")/.."
LIB=visual-slam-pipeline_b200/lib
mkdir -p /tmp/vsm_bench
g++ -O2 -std=c++17 -I include bench_cpp/track_latency.cpp -L $LIB -lvsm -Wl,-rpath,$(pwd)/$LIB -o /tmp/vsm_bench/track_latency
N=${1:-2544}
echo -n '{"tile_top2": '; /tmp/vsm_bench/track_latency $N | tr -d '\n'
echo -n ', "top4": '; VSM_NO_T2=1 /tmp/vsm_bench/track_latency $N | tr -d '\n'
echo -n ', "tile_top2_with_raw_list": '; VSM_TRACK_RAW=1 /tmp/vsm_bench/track_latency $N | tr -d '\n'
echo -n ', "top4_with_raw_list": '; VSM_TRACK_RAW=1 VSM_NO_T2=1 /tmp/vsm_bench/track_latency $N | tr -d '\n'
echo '}'
