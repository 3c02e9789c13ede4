python profiles/ab_bench.py stable,mixed,dense head,main 4 2>&1 | tee gpurun_out/r2h_ab.txt
python -m pytest tests/test_gpu_deepz.py tests/test_gpu_bench_path.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15
