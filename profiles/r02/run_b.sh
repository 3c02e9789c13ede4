set -x
python -m pytest tests/test_gpu_bench_path.py -x -q -m gpu > gpurun_out/r2b_tests.log 2>&1; tail -15 gpurun_out/r2b_tests.log
python profiles/prof_deepz.py stable 300000 2 > gpurun_out/r2b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:deepz_kernel -s 4 -c 1 -f -o gpurun_out/r2b_deepz python profiles/prof_deepz.py stable 300000 2 > gpurun_out/r2b_ncu.log 2>&1
tail -5 gpurun_out/r2b_ncu.log
