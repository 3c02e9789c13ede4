set -x
python profiles/shard_costs.py 8 stable > gpurun_out/r2d_shards.txt 2>&1; cat gpurun_out/r2d_shards.txt
python bench.py --steps 5 --warmup 3 --no-secondary > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -3 gpurun_out/r2d_bench.err; head -c 3000 gpurun_out/r2d_bench.json
