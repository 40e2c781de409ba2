mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --pdl 0 > gpurun_out/r3a_bench_pdl0.json 2> gpurun_out/r3a_bench_pdl0.err; echo "bench pdl0 rc=$?"; cut -c1-330 gpurun_out/r3a_bench_pdl0.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --pdl 1 > gpurun_out/r3a_bench_pdl1.json 2> gpurun_out/r3a_bench_pdl1.err; echo "bench pdl1 rc=$?"; cut -c1-330 gpurun_out/r3a_bench_pdl1.json
tail -3 gpurun_out/r3a_bench_pdl1.err
