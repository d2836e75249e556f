mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err; echo "bench rc=$?"
tail -c 200 gpurun_out/r02_bench9.err
python -c "import __graft_entry__ as g; g.smoke()"
