mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02_pytest4.log 2>&1
tail -4 gpurun_out/r02_pytest4.log
(time python bench.py --config sweep) > gpurun_out/r02_sweep_1gpu.json 2> gpurun_out/r02_sweep_1gpu.err; echo "sweep rc=$?"
tail -c 400 gpurun_out/r02_sweep_1gpu.err
