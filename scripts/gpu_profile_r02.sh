#!/bin/bash
# Run on the GPU box (round 2): full bench line, ncu launch list of one replayed step, ncu --set full of the decoder-cell kernels.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench.json
python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
python scripts/kernel_times.py 2 3 > gpurun_out/plain2.log 2>&1 || { echo "kernel_times failed"; tail -5 gpurun_out/plain2.log; }
for k in fused_cell_fwd_kernel fused_cell_bwd_kernel cell_wgrad_kernel panel_wgrad_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r02_prof_$k \
      python scripts/kernel_times.py 2 3 > gpurun_out/ncu_full_$k.log 2>&1
  echo "ncu full $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
