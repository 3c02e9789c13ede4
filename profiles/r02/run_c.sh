set -x
export CLR_B200_LIB=$PWD/closedloopreachability.jl_b200/libclr_b200_timing.so
for r in stable mixed dense; do python profiles/phase_timing.py $r 200000 0 2>&1 | tail -2; done > gpurun_out/r2c_phase.txt
cat gpurun_out/r2c_phase.txt
unset CLR_B200_LIB
python -m pytest tests/test_solve_order.py -x -q -m gpu 2>&1 | tail -3
