python -m pytest tests/test_gpu_rollout.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15
python profiles/rollout_bench.py 1000000 2>&1 | tail -3
cd tests/c_abi && ls
