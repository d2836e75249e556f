mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --config sweep > gpurun_out/r02_sweep_4gpu.json 2> gpurun_out/r02_sweep_4gpu.err; echo "sweep4 rc=$?"
