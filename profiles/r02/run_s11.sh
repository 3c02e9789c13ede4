python -m pytest tests/test_gpu_bench_path.py tests/test_gpu_deepz.py tests/test_c_abi.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5
python profiles/ab_bench.py stable,mixed,dense main 3 2>&1 | tail -3
python profiles/latency_small.py 2>&1 | tail -3
