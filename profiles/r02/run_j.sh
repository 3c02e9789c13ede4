for w in 0 1; do echo "CLR_B200_WIDE=$w"; CLR_B200_WIDE=$w python profiles/ab_bench.py stable,dense main 4 2>&1; done | tee gpurun_out/r2j_ab.txt
