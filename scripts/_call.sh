timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hub_pieces" > gpurun_out/r2_t19.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2_t19.log
