python -m pytest tests -m gpu -x -q > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7.log
tail -n 30 gpurun_out/r2_t7.log
