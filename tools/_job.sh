mkdir -p gpurun_out
(time python -m pytest tests/test_multi_rank.py -m gpu -q) > gpurun_out/r02_pytest5.log 2>&1
grep -E "^E   .*Assert|^FAILED|passed|failed" gpurun_out/r02_pytest5.log | cut -c1-300
