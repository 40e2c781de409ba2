#!/bin/bash
# GPU box, round-2 final measurement: GPU suite, smoke, full bench line, ncu launch list of one replayed step, graph-build phase
# timeline, ncu --set full of the kernels DESIGN.md quotes.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short --maxfail=8 > gpurun_out/r4h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r4h_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r4h_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r4h_bench.json 2> gpurun_out/r4h_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r4h_bench.json; tail -2 gpurun_out/r4h_bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r4h_bench_reference.json 2> gpurun_out/r4h_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r4h_bench_reference.json
timeout 300 python bench.py --mode infer --steps 32 --warmup 3 > gpurun_out/r4h_bench_infer1.json 2> gpurun_out/r4h_infer.err; echo "infer rc=$?"; cut -c1-300 gpurun_out/r4h_bench_infer1.json
timeout 200 python scripts/build_phases.py > gpurun_out/r4h_graph_build_phases.txt 2>&1; echo "phases rc=$?"
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2400 --csv --log-file gpurun_out/r4h_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 200 python scripts/kernel_times.py 2 3 > gpurun_out/r4h_kernel_times.txt 2>&1 || { echo "kernel_times failed"; tail -5 gpurun_out/r4h_kernel_times.txt; }
for k in fused_cell_fwd_kernel fused_cell_bwd_kernel head_bwd_kernel cell_wgrad_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r4h_prof_$k \
      python scripts/kernel_times.py 2 3 > gpurun_out/ncu_full_$k.log 2>&1
  echo "ncu full $k rc=$?"
done
for k in quadtree_graph_kernel regrid_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r4h_prof_$k \
      python scripts/dynamic_config_times.py > gpurun_out/ncu_full_$k.log 2>&1
  echo "ncu full $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
