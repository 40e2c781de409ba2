mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_graph.py tests/test_golden.py -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/r3o_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3o_pytest.log
timeout 250 python scripts/build_phases.py 2>&1 | tee gpurun_out/r3o_build_phases.log
timeout 250 python scripts/build_profile.py > gpurun_out/r3o_build_profile.log 2>&1; head -4 gpurun_out/r3o_build_profile.log
