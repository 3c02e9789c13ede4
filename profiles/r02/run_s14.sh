python -m pytest tests -x -q -m gpu 2>&1 | tail -4
gcc -O2 -I include profiles/latency_c.c -L closedloopreachability.jl_b200 -lclr_b200 -lm -o /tmp/latency_c
echo "zero copy on:";  LD_LIBRARY_PATH=closedloopreachability.jl_b200 /tmp/latency_c 2>&1 | tail -3
echo "zero copy off:"; CLR_B200_ZEROCOPY=0 LD_LIBRARY_PATH=closedloopreachability.jl_b200 /tmp/latency_c 2>&1 | tail -3
python profiles/latency_small.py 2>&1 | tail -3
