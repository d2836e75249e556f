mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r02_pytest8.log 2>&1
grep -E "^E   |^FAILED|passed|failed" gpurun_out/r02_pytest8.log | cut -c1-300 | head -30
