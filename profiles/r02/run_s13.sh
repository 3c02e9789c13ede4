O=gpurun_out; T=/tmp/ncu_reports; mkdir -p $T
for r in stable dense; do
  s=3; [ $r = dense ] && s=4
  python profiles/prof_deepz.py $r 300000 2 > $O/f2_plain_$r.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:deepz_kernel -s $s -c 1 -f -o $T/deepz_$r python profiles/prof_deepz.py $r 300000 2 > $O/f2_ncu_$r.log 2>&1
  python profiles/ncu_summary.py $T/deepz_$r.ncu-rep deepz_kernel "ncu --set full --clock-control none, main-pass deepz_kernel launch of profiles/prof_deepz.py $r 300000 (a 300 000-cell slice of the bench split)" first > $O/f2_ncu_deepz_$r.json
  ncu -i $T/deepz_$r.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $T/deepz_${r}_src.csv
  python profiles/ncu_lines.py $T/deepz_${r}_src.csv 60 phases > $O/f2_ncu_deepz_${r}_lines.txt
done
