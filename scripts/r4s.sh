#!/bin/bash
# GPU box: full GPU suite + smoke + full bench line of the final build.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short --maxfail=8 > gpurun_out/r4s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r4s_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r4s_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4s_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r4s_bench.json 2> gpurun_out/r4s_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r4s_bench.json
