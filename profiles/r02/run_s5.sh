# ncu of the split rollout kernels (controller_cw_kernel, rollout_ode_kernel)
O=gpurun_out; T=/tmp/ncu_reports; mkdir -p $T
python profiles/rollout_bench.py 1000000 > $O/s5_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:controller_cw -s 2 -c 1 -f -o $T/ctrl python profiles/rollout_bench.py 1000000 > $O/s5_ncu_ctrl.log 2>&1
python profiles/ncu_summary.py $T/ctrl.ncu-rep controller_cw "ncu --set full --clock-control none, controller_cw_kernel (Unicycle 4-500-2, 10^6 points)" first > $O/s5_ncu_ctrl.json
ncu -i $T/ctrl.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/ctrl_src.csv
python profiles/ncu_lines.py $T/ctrl_src.csv 25 > $O/s5_ncu_ctrl_lines.txt
ncu -i $T/ctrl.ncu-rep --page details > $O/s5_ncu_ctrl_details.txt
ncu --set full --clock-control none --import-source on -k regex:rollout_ode -s 2 -c 1 -f -o $T/ode python profiles/rollout_bench.py 1000000 > $O/s5_ncu_ode.log 2>&1
python profiles/ncu_summary.py $T/ode.ncu-rep rollout_ode "ncu --set full --clock-control none, rollout_ode_kernel (Unicycle, Tsit5, one period of 10^6 trajectories)" first > $O/s5_ncu_ode.json
ncu -i $T/ode.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/ode_src.csv
python profiles/ncu_lines.py $T/ode_src.csv 30 > $O/s5_ncu_ode_lines.txt
ncu -i $T/ode.ncu-rep --page details > $O/s5_ncu_ode_details.txt
