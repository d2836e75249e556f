mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02_pytest12.log 2>&1
grep -E "^E   |^FAILED|passed|failed|rror" gpurun_out/r02_pytest12.log | cut -c1-300 | head
python tools/nest_blocks.py 2304 2>&1 | cut -c1-130
python bench.py --steps 5 --warmup 3 --no-secondary > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; echo "bench rc=$?"
