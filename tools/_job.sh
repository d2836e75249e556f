mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r02_pytest16.log 2>&1
grep -E "^E   |^FAILED|passed|failed|rror" gpurun_out/r02_pytest16.log | cut -c1-300 | head -20
