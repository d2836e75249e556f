mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_16M_8gpu.json 2> gpurun_out/r02_bench_16M_8gpu.err; echo "b8 rc=$?"
$T --master-port 29522 bench.py --gpus 8 --config minres3d_diag_cheb --nx 251 --steps 1 --warmup 1 > gpurun_out/r02_bench3d_64M_8gpu.json 2> gpurun_out/r02_bench3d_64M_8gpu.err; echo "3d8 rc=$?"
tail -c 400 gpurun_out/r02_bench3d_64M_8gpu.err
$T --master-port 29523 bench.py --gpus 8 --config fgmres_schur_lsc --nx 96 --steps 2 --warmup 1 > gpurun_out/r02_bench_lsc_nx96_8gpu.json 2> gpurun_out/r02_bench_lsc_nx96_8gpu.err; echo "lsc8 rc=$?"
$T --master-port 29524 bench.py --gpus 8 --config sweep > gpurun_out/r02_sweep_8gpu.json 2> gpurun_out/r02_sweep_8gpu.err; echo "sweep8 rc=$?"
tail -c 300 gpurun_out/r02_sweep_8gpu.err
