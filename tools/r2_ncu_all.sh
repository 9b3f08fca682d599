#!/bin/bash
# ncu --set full of the kernels VERDICT names, one capture each (the first launch = all trials active), headline shape
cd "$(dirname "$0")/.."
REPS=0 bash tools/r2_ncu.sh laplace_eval_kernel eval 0 1
REPS=0 bash tools/r2_ncu.sh pcg_step_kernel pcg 0 1
REPS=0 bash tools/r2_ncu.sh laplace_linesearch_kernel linesearch 0 1
