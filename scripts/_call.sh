python -m pytest tests/test_gpu_patches.py -x -q > gpurun_out/r2_t14.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r2_t14.log
