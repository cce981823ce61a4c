#!/bin/bash
# Manual GPU session (under gpurun): parity tests, then the k_extend variant sweep, then a paths-per-pass sweep.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python tests/gpu_variants.py "$@" > gpurun_out/variants.log 2>&1; tail -20 gpurun_out/variants.log
