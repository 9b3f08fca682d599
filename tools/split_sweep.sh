#!/bin/bash
# sweep of the number of stream parts used for split factorisation / inverse batches (PGPFA_SPLIT)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/split_sweep.txt
for R in 128 1024; do for S in 1 2 3 4; do
  PGPFA_SPLIT=$S timeout 300 python bench.py --trials $R --skip-cpu --steps 4 --warmup 3 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('R=$R split=$S', round(d['ms_per_step'],2), 'ms chol', round(r['achieved'],2), 'TF trtri', round(r['trtri_tflops'],2), {k:round(v,2) for k,v in r['other_ms_per_step'].items()})" >> gpurun_out/split_sweep.txt
done; done
cat gpurun_out/split_sweep.txt
