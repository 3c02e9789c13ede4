python -m pytest tests/test_gpu_deepz.py -x -q -m gpu -k "reduction" 2>&1 | tail -5
python -m pytest tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -3
echo "--- staged + register sort"; python profiles/reduce_bench.py 6 2>&1 | tail -2
echo "--- shared-memory sort (round 2 session 1)"; CLR_B200_REDUCE_REG=0 python profiles/reduce_bench.py 6 2>&1 | tail -2
