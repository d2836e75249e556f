mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r02_pytest13.log 2>&1
grep -E "^E   |^FAILED|passed|failed|rror" gpurun_out/r02_pytest13.log | cut -c1-300 | head -20
python -c "import __graft_entry__ as g; g.smoke()"
