#!/bin/bash
# ncu --set full of the kernels VERDICT names, one capture each, in a warm E-step / EM iteration at the headline shape
cd "$(dirname "$0")/.."
bash tools/r2_ncu.sh laplace_eval_kernel eval 6 1
bash tools/r2_ncu.sh pcg_step_kernel pcg 60 1
bash tools/r2_ncu.sh prior_apply_kernel prior 100 1
bash tools/r2_ncu.sh syrk_sum_kernel syrk 1 1
MODE=emstep bash tools/r2_ncu.sh mstep_cd_stats_kernel cdstats 5 1
