#!/bin/bash
# ncu --set full of the hot kernels, one capture each, at the headline shape (tools/prof_lowrank.py); the reports are
# summarised on the box (tools/ncu_summary.py) and deleted: only the text summaries travel back.
# REPS=0: the first E-step, i.e. launch 0 of the CG / evaluation kernels has all 1024 trials active.
cd "$(dirname "$0")/.."
cap() {   # regex name skip count [env]
    bash tools/r2_ncu.sh "$1" "$2" "$3" "$4" > /dev/null 2>&1
    { echo "# ncu --set full --clock-control none --import-source on -k regex:$1 -s $3 -c $4 python tools/prof_lowrank.py (MODE=${MODE:-estep} REPS=${REPS:-1}; 1024 trials, q=8, N=100, T=200, r=179)";
      python tools/ncu_summary.py gpurun_out/prof_$2.ncu-rep 16; } > gpurun_out/ncu_$2.txt 2>&1
    rm -f gpurun_out/prof_$2.ncu-rep gpurun_out/ncu_$2.log
}
export REPS=0
cap laplace_eval_kernel eval 0 1
cap pcg_step_kernel pcg 0 1
cap prior_apply_kernel prior 2 1      # launch 2 = the first stacked M^-1 / K^-1 M^-1 apply
cap laplace_linesearch_kernel linesearch 0 1
cap syrk_sum_kernel syrk 0 1
cap lr_mix_kernel mix 0 1
cap lr_vsm_kernel vsm 0 1
cap spd_sweep_kernel sweep 0 1
cap gemm_nt_kernel gemm 0 2
MODE=emstep REPS=1 cap mstep_cd_stats_kernel cdstats 5 1
ls -la gpurun_out/
