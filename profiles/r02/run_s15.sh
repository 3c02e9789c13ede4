# compute-sanitizer memcheck over the kernels touched in the last session (small problem sizes)
export CLR_FUZZ_TRIALS=6
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "narrow or constant_bank or split_rollout or step_log or other_plants or runtime_compiled" 2>&1 | tail -6
echo "rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_deepz.py -x -q -m gpu -k "reduction or tora or acc or fuzz or quadrotor or vertical" 2>&1 | tail -6
echo "rc=$?"
