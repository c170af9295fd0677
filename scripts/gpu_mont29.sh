#!/bin/bash
# round 2: whole-product throughput of the 4-limb Montgomery formulations (scripts/mont29_bench.cu) + ncu of a few
set -u
mkdir -p gpurun_out
cd scripts
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../thaler_study_b200/csrc -I../include -o /tmp/mont29_bench mont29_bench.cu 2> ../gpurun_out/mont29_build.err
timeout 300 /tmp/mont29_bench > ../gpurun_out/mont29_bench.jsonl 2> ../gpurun_out/mont29_bench.err
grep probe ../gpurun_out/mont29_bench.jsonl | cut -c1-330
timeout 600 ncu --set full --clock-control none --csv --page raw --log-file ../gpurun_out/mont29_ncu_raw.csv /tmp/mont29_bench prof > ../gpurun_out/mont29_ncu.log 2>&1
tail -3 ../gpurun_out/mont29_ncu.log
