#!/bin/bash
# Run on the GPU box: bench (plain), then the ncu launch list of the same command, then one ncu --set full capture of
# the roofline kernel (decoder cell forward).
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
rc=$?
echo "bench rc=$rc"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ $rc -eq 0 ]; then
  python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 4500 -c 4400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
  echo "ncu launches rc=$?"; tail -2 gpurun_out/ncu_launch.log
  python scripts/kernel_times.py 2 3 > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:fused_cell_fwd_kernel -s 4 -c 1 -o gpurun_out/prof_cell \
      python scripts/kernel_times.py 2 3 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
