#!/bin/bash
# round-2 GPU call: new device-loop tests (full traceback), the whole GPU suite, smoke, parity table, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_device_loops.py -m gpu -x -q --timeout 300 2>&1 | tail -60 ) > gpurun_out/loops.log; tail -30 gpurun_out/loops.log
( timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 ) > gpurun_out/tests.log; tail -15 gpurun_out/tests.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log; cat gpurun_out/smoke.log
if [ -n "$PARITY" ]; then ( timeout 900 python tools/parity_report.py 2> gpurun_out/parity.err ) > gpurun_out/parity_report.json; cat gpurun_out/parity_report.json; tail -5 gpurun_out/parity.err; fi
if [ -z "$NOBENCH" ]; then ( timeout 1500 python bench.py "$@" 2> gpurun_out/bench.err | tail -3 ) > gpurun_out/bench.json; cat gpurun_out/bench.json; tail -20 gpurun_out/bench.err; fi
