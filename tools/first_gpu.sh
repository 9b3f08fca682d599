#!/bin/bash
# first GPU contact: run kernel parity tests, keep full log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/first_gpu.log
cat gpurun_out/first_gpu.log
