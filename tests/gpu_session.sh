#!/bin/bash
# Manual GPU session (under gpurun): parity tests, then the k_extend build-variant sweep (args before "--"),
# then an environment sweep with the default build (args after "--").
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
variants=(); envs=(); cur=v
for a in "$@"; do if [ "$a" == "--" ]; then cur=e; elif [ $cur == v ]; then variants+=("$a"); else envs+=("$a"); fi; done
if [ ${#variants[@]} -gt 0 ]; then python tests/gpu_variants.py "${variants[@]}" > gpurun_out/variants.log 2>&1; tail -20 gpurun_out/variants.log; fi
if [ ${#envs[@]} -gt 0 ]; then python tests/gpu_env_sweep.py "${envs[@]}" > gpurun_out/env_sweep.log 2>&1; tail -20 gpurun_out/env_sweep.log; fi
