mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rollout_pool.py tests/test_parity_graph.py tests/test_golden.py -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/r3i_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r3i_pytest.log
timeout 300 python bench.py --mode infer --steps 32 --warmup 3 > gpurun_out/r3i_infer1.json 2> gpurun_out/r3i_infer1.err; echo "infer rc=$?"; cut -c1-400 gpurun_out/r3i_infer1.json; tail -3 gpurun_out/r3i_infer1.err
timeout 300 python bench.py --mode infer --steps 32 --warmup 3 --lanes 1 > gpurun_out/r3i_infer1_lane1.json 2> gpurun_out/r3i_infer1_lane1.err; echo "infer lanes1 rc=$?"; cut -c1-200 gpurun_out/r3i_infer1_lane1.json
