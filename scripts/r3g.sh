mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/r3g_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3g_pytest.log
timeout 250 python scripts/build_profile.py > gpurun_out/r3g_build_profile.log 2>&1; head -4 gpurun_out/r3g_build_profile.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --pdl 1 > gpurun_out/r3g_bench_pdl1.json 2> gpurun_out/r3g_bench_pdl1.err; echo "bench pdl1 rc=$?"; cut -c1-330 gpurun_out/r3g_bench_pdl1.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --pdl 0 > gpurun_out/r3g_bench_pdl0.json 2> gpurun_out/r3g_bench_pdl0.err; echo "bench pdl0 rc=$?"; cut -c1-330 gpurun_out/r3g_bench_pdl0.json
