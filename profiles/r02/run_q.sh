python -m pytest tests/test_gpu_deepz.py -x -q -m gpu 2>&1 | tail -12
python profiles/latency_small.py 2>&1 | tail -4
