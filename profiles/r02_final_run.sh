# Final measurements of round 2 on one B200: bench lines (both arms), small-call latencies, launch list, ncu --set full
# summaries (processed on the box: the reports are too large to travel).  Usage: gpurun -- bash profiles/r02_final_run.sh
set -x
O=gpurun_out
python bench.py > $O/f2_bench.json 2> $O/f2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/f2_bench_ref.json 2>> $O/f2_bench.err
python profiles/latency_small.py > $O/f2_latency.txt 2>&1
(gcc -O2 -I include profiles/latency_c.c -L closedloopreachability.jl_b200 -lclr_b200 -lm -o /tmp/latency_c && LD_LIBRARY_PATH=closedloopreachability.jl_b200 /tmp/latency_c > $O/f2_latency_c.txt 2>&1) || true
python profiles/reduce_bench.py 6 > $O/f2_reduce.txt 2>&1
python profiles/project_bench.py > $O/f2_project.txt 2>&1
python bench.py --steps 2 --warmup 1 --no-regimes > $O/f2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/f2_launches.csv python bench.py --steps 2 --warmup 1 --no-regimes > $O/f2_ncu_launch.log 2>&1
T=/tmp/ncu_reports; mkdir -p $T
for r in stable mixed dense; do
  python profiles/prof_deepz.py $r 300000 2 > $O/f2_plain_$r.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:deepz_kernel -s $([ $r = dense ] && echo 4 || echo 3) -c 1 -f -o $T/deepz_$r python profiles/prof_deepz.py $r 300000 2 > $O/f2_ncu_$r.log 2>&1
  python profiles/ncu_summary.py $T/deepz_$r.ncu-rep deepz_kernel "ncu --set full --clock-control none, main-pass deepz_kernel launch of profiles/prof_deepz.py $r 300000 (a 300 000-cell slice of the bench split)" first > $O/f2_ncu_deepz_$r.json
  ncu -i $T/deepz_$r.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/deepz_${r}_src.csv
  python profiles/ncu_lines.py $T/deepz_${r}_src.csv 40 phases > $O/f2_ncu_deepz_${r}_lines.txt
done
python profiles/quad_passes.py > $O/f2_plain_quad.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:deepz_kernel -s 2 -c 1 -f -o $T/quad python profiles/quad_passes.py > $O/f2_ncu_quad.log 2>&1
python profiles/ncu_summary.py $T/quad.ncu-rep deepz_kernel "ncu --set full --clock-control none, deepz_kernel of one C3 call (Quadrotor, 4096 cells, two cells per tile)" first > $O/f2_ncu_quad.json
ncu -i $T/quad.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/quad_src.csv
python profiles/ncu_lines.py $T/quad_src.csv 30 phases > $O/f2_ncu_quad_lines.txt
python profiles/rollout_bench.py 1000000 > $O/f2_plain_rollout.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:controller_cw -s 2 -c 1 -f -o $T/ctrl python profiles/rollout_bench.py 1000000 > $O/f2_ncu_ctrl.log 2>&1
python profiles/ncu_summary.py $T/ctrl.ncu-rep controller_cw "ncu --set full --clock-control none, controller_cw_kernel (Unicycle 4-500-2, one control period of 10^6 trajectories)" first > $O/f2_ncu_controller_cw.json
ncu -i $T/ctrl.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/ctrl_src.csv
python profiles/ncu_lines.py $T/ctrl_src.csv 20 > $O/f2_ncu_controller_cw_lines.txt
ncu --set full --clock-control none --import-source on -k regex:rollout_ode -s 2 -c 1 -f -o $T/ode python profiles/rollout_bench.py 1000000 > $O/f2_ncu_ode.log 2>&1
python profiles/ncu_summary.py $T/ode.ncu-rep rollout_ode "ncu --set full --clock-control none, rollout_ode_kernel (Unicycle, Tsit5, one control period of 10^6 trajectories)" first > $O/f2_ncu_rollout_ode.json
ncu -i $T/ode.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/ode_src.csv
python profiles/ncu_lines.py $T/ode_src.csv 25 > $O/f2_ncu_rollout_ode_lines.txt
python profiles/reduce_bench.py 5 > $O/f2_plain_reduce.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:reduce_order -c 1 -f -o $T/reduce python profiles/reduce_bench.py 5 > $O/f2_ncu_reduce.log 2>&1
python profiles/ncu_summary.py $T/reduce.ncu-rep reduce_order "ncu --set full --clock-control none, reduce_order_reg_kernel<8> (Quadrotor, 5^6 cells, order 2)" first > $O/f2_ncu_reduce.json
ncu -i $T/reduce.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/reduce_src.csv
python profiles/ncu_lines.py $T/reduce_src.csv 15 > $O/f2_ncu_reduce_lines.txt
du -sh $O
