python -m pytest tests -m gpu -x -q > gpurun_out/r2_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t6.log
tail -n 30 gpurun_out/r2_t6.log
python scripts/bwd_timing.py 64 256 64
python scripts/bwd_timing.py 16 256 32
