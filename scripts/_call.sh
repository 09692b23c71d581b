python -m pytest tests -m gpu -x -q > gpurun_out/r2_t12.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r2_t12.log
python scripts/patch_config4.py --iters 5 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg4.csv python scripts/patch_config4.py --iters 1 > gpurun_out/ncu_cfg4.log 2>&1
tail -n 2 gpurun_out/plain.log
python bench.py --steps 500 --no-also > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2_bench3.err
python bench.py --workload generator --steps 10 > gpurun_out/r2_gen2.json 2> gpurun_out/r2_gen2.err; echo "gen rc=$?"; tail -n 5 gpurun_out/r2_gen2.err
