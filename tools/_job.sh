mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x) > gpurun_out/r02_pytest15.log 2>&1
grep -E "^E   |^FAILED|passed|failed|rror" gpurun_out/r02_pytest15.log | cut -c1-300 | head -20
(time python bench.py --config sweep) > gpurun_out/r02_sweep_1gpu.json 2> gpurun_out/r02_sweep_1gpu.err; echo "sweep rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench rc=$?"
