#!/bin/bash
# standard GPU round: tests + smoke + bench (logs in gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 ) > gpurun_out/tests.log
tail -5 gpurun_out/tests.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log; cat gpurun_out/smoke.log
( timeout 1500 python bench.py "$@" 2> gpurun_out/bench.err | tail -3 ) > gpurun_out/bench.json; cat gpurun_out/bench.json; tail -20 gpurun_out/bench.err
