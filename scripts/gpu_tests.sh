#!/bin/bash
# Run on the GPU box: full GPU parity suite (stops after 12 failures) + smoke.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short --maxfail=12 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -40 gpurun_out/pytest_gpu.log
tail -5 gpurun_out/smoke.log
