# the driver's round-end sequence on the final tree: build check is done on the CPU side; here smoke, the GPU suite, both bench arms
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/final_bench.json | head -c 400; echo
python bench.py --impl reference --steps 3 --warmup 1 2> /dev/null | head -c 300; echo
