mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/r3p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3p_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r3p_bench.json 2> gpurun_out/r3p_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r3p_bench.json; tail -3 gpurun_out/r3p_bench.err
