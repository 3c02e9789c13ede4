python profiles/ab_bench.py stable,mixed,dense head,main,head,main 4 2>&1 | tee gpurun_out/r2k_ab.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
