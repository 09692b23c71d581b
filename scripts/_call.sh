timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t18.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2_t18.log
python bench.py --steps 500 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"
