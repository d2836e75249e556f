mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_spmv_pd -c 1 -s 5 -o gpurun_out/r02_ncu_spmv_pd_A_v2 python tools/spmv_only.py 2304 4 A > gpurun_out/r02_ncu_A2.log 2>&1; echo "ncu full rc=$?"
