O=gpurun_out; T=/tmp/ncu_reports; mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:reduce_order -c 1 -f -o $T/reduce python profiles/reduce_bench.py 5 > $O/s9_ncu_reduce.log 2>&1
python profiles/ncu_summary.py $T/reduce.ncu-rep reduce_order "ncu --set full --clock-control none, reduce_order_reg_kernel<8> (Quadrotor, 5^6 cells, order 2, generators staged in shared memory, keys sorted in registers)" first > $O/s9_ncu_reduce.json
ncu -i $T/reduce.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/reduce_src.csv
python profiles/ncu_lines.py $T/reduce_src.csv 25 > $O/s9_ncu_reduce_lines.txt
ncu -i $T/reduce.ncu-rep --page details | grep -E 'Occupancy|Ipc|Eligible|Issued|Registers|Shared Memory|Bank' > $O/s9_ncu_reduce_details.txt
