#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --profile-mode \
    > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_bench.err
wc -l gpurun_out/launches_bench.csv
