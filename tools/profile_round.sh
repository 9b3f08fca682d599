#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# launch list of the bench command (every kernel: device time + DRAM bytes; cold-cache & serialised: compare shares)
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --profile-mode \
    > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_bench.err
wc -l gpurun_out/launches_bench.csv
# full capture of the dominant kernel of the low-rank posterior pass (the three gemm_nt launches of a warm E-step)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 3 -c 3 \
    -o gpurun_out/prof_gemm_nt -f python tools/prof_lowrank.py > gpurun_out/ncu_gemm.log 2>&1
# dense path (DENSE=1): panel kernel, two mid-factorisation launches of the 1024-trial factorisation
if [ -n "$DENSE" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_panel_kernel -s 15 -c 2 \
    -o gpurun_out/prof_panel -f python tools/prof_factor.py > gpurun_out/ncu_panel.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
