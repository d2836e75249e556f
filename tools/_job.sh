mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_16M_8gpu.json 2> gpurun_out/r02_bench_16M_8gpu.err; echo "b8 rc=$?"
