#!/bin/bash
# multi-GPU check: the world-size test (N=2 only) and the headline bench line at N ranks
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then ( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 500 2>&1 | tail -4 ) > gpurun_out/multi_test.log; cat gpurun_out/multi_test.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --headline-only --skip-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.json | cut -c1-330; tail -3 gpurun_out/bench_n$N.err
