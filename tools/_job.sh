mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-secondary > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err; echo "bench rc=$?"
