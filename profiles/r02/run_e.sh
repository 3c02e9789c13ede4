set -x
python -m pytest tests/test_gpu_deepz.py tests/test_gpu_bench_path.py -x -q -m gpu > gpurun_out/r2e_tests.log 2>&1; tail -12 gpurun_out/r2e_tests.log
python profiles/ab_bench.py stable,mixed,dense main 5 2>&1 | tee gpurun_out/r2e_ab.txt
python profiles/shard_costs.py 8 stable > gpurun_out/r2e_shards.txt 2>&1; cat gpurun_out/r2e_shards.txt
python bench.py --steps 5 --warmup 3 --no-secondary > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err; head -c 1500 gpurun_out/r2e_bench.json
