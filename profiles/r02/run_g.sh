for v in headT curT; do
  export CLR_B200_LIB=$PWD/closedloopreachability.jl_b200/libclr_b200_$v.so
  for r in stable dense; do echo "$v $r: $(python profiles/phase_timing.py $r 200000 0 2>&1 | tail -1)"; done
done | tee gpurun_out/r2g_phase.txt
