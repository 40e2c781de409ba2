#!/bin/bash
# GPU box: GPU parity suite, smoke, the full bench line, the dynamic-mesh configurations.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short --maxfail=8 > gpurun_out/r4a_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r4a_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r4a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r4a_bench.json; tail -3 gpurun_out/r4a_bench.err
timeout 300 python scripts/dynamic_config_times.py > gpurun_out/r4a_dynamic.txt 2>&1; echo "dyn rc=$?"; tail -4 gpurun_out/r4a_dynamic.txt
