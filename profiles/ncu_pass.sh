# usage: bash profiles/ncu_pass.sh <tag> <skip> <driver command...>   -- ncu --set full of the (skip+1)-th deepz_kernel launch of the command,
# summarised on the box (the reports are too large to travel): gpurun_out/<tag>.json, gpurun_out/<tag>_lines.txt
tag=$1; skip=$2; shift 2
O=gpurun_out; T=/tmp/ncu_reports; mkdir -p $T $O
"$@" > $O/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/${tag}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:deepz_kernel -s $skip -c 1 -f -o $T/$tag "$@" > $O/${tag}_ncu.log 2>&1
python profiles/ncu_summary.py $T/$tag.ncu-rep deepz_kernel "ncu --set full --clock-control none, launch $skip of: $*" first > $O/$tag.json
ncu -i $T/$tag.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/${tag}_src.csv
python profiles/ncu_lines.py $T/${tag}_src.csv 45 phases > $O/${tag}_lines.txt
