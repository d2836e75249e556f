mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02_pytest3.log 2>&1
tail -5 gpurun_out/r02_pytest3.log
python bench.py --steps 5 --warmup 3 --no-secondary > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_spmv_pd -c 1 -s 5 -o gpurun_out/r02_ncu_spmv_pd_A python tools/spmv_only.py 2304 4 A > gpurun_out/r02_ncu_A.log 2>&1
tail -2 gpurun_out/r02_ncu_A.log
