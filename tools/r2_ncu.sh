#!/bin/bash
# ncu --set full of one kernel (regex $1) in a warm E-step of the headline shape; report in gpurun_out/prof_$2.ncu-rep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-1} -c ${4:-1} \
    -o gpurun_out/prof_$2 -f python tools/prof_lowrank.py > gpurun_out/ncu_$2.log 2>&1
tail -3 gpurun_out/ncu_$2.log
ls -la gpurun_out/prof_$2.ncu-rep
