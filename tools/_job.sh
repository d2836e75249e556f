mkdir -p gpurun_out
export B200SP_PD_DEBUG=1
(time python -m pytest tests -m gpu -q) > gpurun_out/r02_pytest17.log 2>&1
grep -E "^E   |^FAILED|passed|failed|rror" gpurun_out/r02_pytest17.log | cut -c1-300 | head -20
python bench.py --config minres3d_diag_cheb --nx 96 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench3d_nx96_1gpu_pd.json 2> gpurun_out/r02_bench3d_nx96_1gpu_pd.err; echo "3d rc=$?"
grep "b200sp pd" gpurun_out/r02_bench3d_nx96_1gpu_pd.err | head; tail -c 300 gpurun_out/r02_bench3d_nx96_1gpu_pd.err
