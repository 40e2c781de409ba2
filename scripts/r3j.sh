mkdir -p gpurun_out
timeout 300 python bench.py --mode infer --steps 32 --warmup 3 > gpurun_out/r3j_infer1.json 2> gpurun_out/r3j_infer1.err; echo "infer rc=$?"; cut -c1-250 gpurun_out/r3j_infer1.json; tail -3 gpurun_out/r3j_infer1.err
timeout 300 python bench.py --mode infer --steps 32 --warmup 3 --lanes 1 > gpurun_out/r3j_infer1_lane1.json 2> gpurun_out/r3j_infer1_lane1.err; echo "infer lanes1 rc=$?"; cut -c1-200 gpurun_out/r3j_infer1_lane1.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r3j_bench.json; tail -3 gpurun_out/r3j_bench.err
