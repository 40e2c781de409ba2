#!/bin/bash
# Run on the GPU box: bench (plain), then the ncu launch list of the same command.
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
rc=$?
echo "bench rc=$rc"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 6000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
  echo "ncu rc=$?"; tail -2 gpurun_out/ncu_launch.log
fi
