mkdir -p gpurun_out
python bench.py --config fgmres_schur_lsc --nx 96 --steps 3 --warmup 2 --no-secondary > gpurun_out/r02_bench_lsc_nx96_1gpu.json 2> gpurun_out/r02_bench_lsc_nx96_1gpu.err; echo "lsc rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config fgmres_schur_lsc --nx 96 --steps 3 --warmup 2 > gpurun_out/r02_bench_lsc_nx96_2gpu.json 2> gpurun_out/r02_bench_lsc_nx96_2gpu.err; echo "lsc2 rc=$?"
tail -c 600 gpurun_out/r02_bench_lsc_nx96_2gpu.err
