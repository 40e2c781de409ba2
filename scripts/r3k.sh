mkdir -p gpurun_out
for L in 2 6 8 12; do
timeout 300 python bench.py --mode infer --steps 48 --warmup 3 --lanes $L > gpurun_out/r3k_infer_l$L.json 2> gpurun_out/r3k_infer_l$L.err; echo "lanes $L rc=$?"; cut -c1-140 gpurun_out/r3k_infer_l$L.json
done
