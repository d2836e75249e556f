mkdir -p gpurun_out
(time python -m pytest tests/test_3d.py -m gpu -q) > gpurun_out/r02_pytest9.log 2>&1
grep -E "^E   |^FAILED|passed|failed|rror" gpurun_out/r02_pytest9.log | cut -c1-400 | head -30
python bench.py --config minres3d_diag_cheb --nx 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench3d_nx128.json 2> gpurun_out/r02_bench3d_nx128.err; echo "3d128 rc=$?"
tail -c 300 gpurun_out/r02_bench3d_nx128.err
