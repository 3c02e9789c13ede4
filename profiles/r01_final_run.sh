# Final measurements of a round on one B200: bench lines, launch list, ncu --set full summaries (processed on the box:
# the reports are too large to travel).  Usage: gpurun -- bash profiles/r01_final_run.sh
set -x
O=gpurun_out
python bench.py > $O/f_bench_stable.json 2> $O/f_bench.err
python bench.py --regime mixed --steps 5 > $O/f_bench_mixed.json 2>> $O/f_bench.err
python bench.py --regime dense --steps 3 > $O/f_bench_dense.json 2>> $O/f_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/f_bench_ref.json 2>> $O/f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/f_launches.csv python bench.py --steps 2 --warmup 1 > $O/f_ncu_launch.log 2>&1
T=/tmp/ncu_reports; mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:deepz_kernel -c 6 -o $T/deepz python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $O/f_ncu_deepz.log 2>&1
python profiles/ncu_summary.py $T/deepz.ncu-rep deepz_kernel "ncu --set full --clock-control none, main-pass deepz_kernel launch of bench.py (10^6 cells, stable regime)" longest "<25, 0, 0, 0>" > $O/f_ncu_deepz.json
ncu -i $T/deepz.ncu-rep --page raw --csv --metrics gpu__time_duration.sum > $O/f_ncu_deepz_launches.csv
ncu -i $T/deepz.ncu-rep --page source --csv --print-source cuda,sass --launch-skip 4 --launch-count 1 > $T/deepz_src.csv
python profiles/ncu_lines.py $T/deepz_src.csv 60 > $O/f_ncu_deepz_lines.txt
ncu --set full --clock-control none --import-source on -k regex:rollout_blocked -c 2 -o $T/rollout python profiles/rollout_bench.py 200000 > $O/f_ncu_rollout.log 2>&1
python profiles/ncu_summary.py $T/rollout.ncu-rep rollout_blocked "ncu --set full --clock-control none, rollout_blocked_kernel (Unicycle, 200000 trajectories x 50 periods, Tsit5)" first > $O/f_ncu_rollout.json
ncu -i $T/rollout.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/rollout_src.csv
python profiles/ncu_lines.py $T/rollout_src.csv 30 > $O/f_ncu_rollout_lines.txt
ncu --set full --clock-control none --import-source on -k regex:reduce_order -c 2 -o $T/reduce python profiles/reduce_bench.py 5 > $O/f_ncu_reduce.log 2>&1
python profiles/ncu_summary.py $T/reduce.ncu-rep reduce_order "ncu --set full --clock-control none, reduce_order_kernel (Quadrotor, 5^6 cells, order 2)" first > $O/f_ncu_reduce.json
ncu -i $T/reduce.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/reduce_src.csv
python profiles/ncu_lines.py $T/reduce_src.csv 20 > $O/f_ncu_reduce_lines.txt
du -sh $O
