mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r02_pytest14.log 2>&1
grep -E "^E   |^FAILED|passed|failed|rror" gpurun_out/r02_pytest14.log | cut -c1-300 | head -20
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02_bench7.err
