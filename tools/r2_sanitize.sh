#!/bin/bash
# compute-sanitizer over two whole EM iterations at a reduced trial count (every round-2 kernel runs: low-rank pass with
# the big capacitance product, SYRK/PautoSum, sweep inverse, stacked prior apply + fused CG step, device-driven M-step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${R:-48} MODE=emstep REPS=1 timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prof_lowrank.py > gpurun_out/memcheck_r02.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_r02.txt
tail -4 gpurun_out/memcheck_r02.txt
R=${R:-48} MODE=emstep REPS=0 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prof_lowrank.py > gpurun_out/racecheck_r02.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck_r02.txt
tail -4 gpurun_out/racecheck_r02.txt
